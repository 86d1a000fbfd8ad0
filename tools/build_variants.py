"""Build kernel-variant libraries under epidemicmodeling_b200/variants/ (git-ignored, ships to the GPU box).
    python tools/build_variants.py name:-DFLAG=1,-DOTHER=2 name2:...
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epidemicmodeling_b200 import _build

for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    d = os.path.join(_build.PKG, "variants", name)
    os.makedirs(d, exist_ok=True)
    lib = _build.build(force=True, extra_flags=[f for f in flags.split(",") if f], lib=os.path.join(d, "libepi_b200.so"),
                       obj_dir=os.path.join(d, "build"))
    log = open(os.path.join(d, "build", "eks_gain.o.log")).read()
    for f in os.listdir(os.path.join(d, "build")):   # the objects are not needed on the GPU box (snapshot size limit)
        if f.endswith(".o"):
            os.remove(os.path.join(d, "build", f))
    import re
    for m in re.finditer(r"Function properties for (\S*eks_gain_kernelILi2ELb1E\S*)\n\s*(.*)\nptxas info\s*: Used (\d+) registers", log):
        print(name, m.group(2).strip(), "regs", m.group(3))
