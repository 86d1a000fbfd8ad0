// mex_driver.cpp -- runs matlab/epi_mex.cpp's mexFunction on arguments read from a file and writes the
// outputs (or the error id/message) to another: `mex_driver in.bin out.bin`.  The Python side is
// tests/mexhost/__init__.py.  Stream format (little endian), one node per array:
//   int32 tag (0 double, 1 char, 2 struct, 3 logical, 4 null) | int32 ndims | int64 dims[ndims] | payload
//   double: prod(dims) float64, column-major     char: prod(dims) bytes     logical: prod(dims) bytes
//   struct: int32 nfields, then per field (int32 len, bytes), then per element per field one node
// File: int32 nlhs, int32 nrhs, nrhs nodes.   Reply: int32 n (>= 0: n nodes; -1: int32 len,id,int32 len,msg).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "mex_emul.h"

extern "C" void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

static void rd(FILE *f, void *p, size_t n) {
  if (n && fread(p, 1, n, f) != n) throw std::runtime_error("short read");
}
static int32_t rd32(FILE *f) { int32_t v; rd(f, &v, 4); return v; }
static mxArray *read_node(FILE *f) {
  const int32_t tag = rd32(f);
  if (tag == 4) return nullptr;
  const int32_t nd = rd32(f);
  std::vector<mwSize> dims((size_t)nd);
  size_t n = 1;
  for (auto &d : dims) { int64_t v; rd(f, &v, 8); d = (mwSize)v; n *= d; }
  if (tag == 0) {
    mxArray *a = mxCreateNumericArray((mwSize)nd, dims.data(), mxDOUBLE_CLASS, mxREAL);
    rd(f, mxGetPr(a), n * 8);
    return a;
  }
  if (tag == 1) {
    std::string s(n, ' ');
    rd(f, &s[0], n);
    return mex_emul_char(s);
  }
  if (tag == 3) {
    mxArray *a = mxCreateLogicalMatrix(dims[0], n / (dims[0] ? dims[0] : 1));
    std::vector<unsigned char> b(n);
    rd(f, b.data(), n);
    for (size_t i = 0; i < n; ++i) mxGetLogicals(a)[i] = b[i] != 0;
    return a;
  }
  if (tag == 2) {
    const int32_t nf = rd32(f);
    std::vector<std::string> names((size_t)nf);
    std::vector<const char *> cn;
    for (auto &s : names) { const int32_t l = rd32(f); s.assign((size_t)l, ' '); rd(f, &s[0], (size_t)l); }
    for (auto &s : names) cn.push_back(s.c_str());
    mxArray *a = mxCreateStructMatrix(dims[0], n / (dims[0] ? dims[0] : 1), nf, cn.data());
    for (size_t i = 0; i < n; ++i)
      for (int32_t k = 0; k < nf; ++k) mxSetField(a, i, cn[(size_t)k], read_node(f));
    return a;
  }
  throw std::runtime_error("bad tag");
}
static void wr(FILE *f, const void *p, size_t n) { fwrite(p, 1, n, f); }
static void wr32(FILE *f, int32_t v) { wr(f, &v, 4); }
static void write_node(FILE *f, const mxArray *a) {
  if (!a) { wr32(f, 4); return; }
  const int32_t tag = a->cls == mxDOUBLE_CLASS ? 0 : a->cls == mxCHAR_CLASS ? 1 : a->cls == mxLOGICAL_CLASS ? 3 : 2;
  wr32(f, tag);
  wr32(f, (int32_t)a->dims.size());
  size_t n = 1;
  for (mwSize d : a->dims) { const int64_t v = (int64_t)d; wr(f, &v, 8); n *= d; }
  if (tag == 0) wr(f, a->re.data(), n * 8);
  else if (tag == 1) wr(f, a->chars.data(), n);
  else if (tag == 3) for (size_t i = 0; i < n; ++i) { const unsigned char b = a->lg[i] ? 1 : 0; wr(f, &b, 1); }
  else throw std::runtime_error("struct outputs are not used by epi_mex");
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: mex_driver in.bin out.bin\n"); return 2; }
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) { perror(argv[1]); return 2; }
  std::vector<mxArray *> prhs;
  int nlhs = 0;
  try {
    nlhs = rd32(fi);
    const int nrhs = rd32(fi);
    for (int i = 0; i < nrhs; ++i) prhs.push_back(read_node(fi));
  } catch (const std::exception &e) { fprintf(stderr, "mex_driver: %s\n", e.what()); return 2; }
  fclose(fi);
  std::vector<mxArray *> plhs((size_t)(nlhs > 0 ? nlhs : 1) + 16, nullptr);
  FILE *fo = fopen(argv[2], "wb");
  if (!fo) { perror(argv[2]); return 2; }
  try {
    mexFunction(nlhs, plhs.data(), (int)prhs.size(), const_cast<const mxArray **>(prhs.data()));
    const int n = nlhs > 0 ? nlhs : 1;
    wr32(fo, n);
    for (int i = 0; i < n; ++i) write_node(fo, plhs[(size_t)i]);
  } catch (const MexError &e) {
    wr32(fo, -1);
    wr32(fo, (int32_t)e.id.size()); wr(fo, e.id.data(), e.id.size());
    wr32(fo, (int32_t)e.msg.size()); wr(fo, e.msg.data(), e.msg.size());
  }
  fclose(fo);
  mex_emul_run_at_exit();
  return 0;
}
