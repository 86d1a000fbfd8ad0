# piped schedule knobs at the 8-GPU shard: EPI_PIPE_WARPS x EPI_PIPE_CHUNKS
for W in ${WARPS:-2 3 4}; do for C in ${CHUNKS:-8}; do
EPI_PIPE_WARPS=$W EPI_PIPE_CHUNKS=$C timeout 60 python bench.py --regions ${REG:-30} --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/ch_${W}_$C.log 2> gpurun_out/ch_${W}_$C.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ch_${W}_$C.log").read().strip().splitlines()[-1])
    print("warps $W chunks $C ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("warps $W chunks $C FAILED", e)
PY
done; done
