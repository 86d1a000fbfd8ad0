// mex_emul.cpp -- a small in-memory implementation of the MEX C API subset declared in tests/stubs/mex.h,
// so that matlab/epi_mex.cpp can be LINKED AND EXECUTED without MATLAB/Octave (tests/mexhost/mex_driver.cpp
// feeds mexFunction from a file and writes its outputs back).  Semantics follow the documented MEX API:
// column-major numeric arrays, struct arrays addressed (element, field), 1-based nothing (indices are 0-based
// in the C API), mexErrMsgIdAndTxt never returns.  Test infrastructure only.
#include "mex_emul.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

static size_t numel(const mxArray *a) {
  size_t n = 1;
  for (mwSize d : a->dims) n *= d;
  return n;
}
static mxArray *make(mxClassID cls, std::vector<mwSize> dims) {
  mxArray *a = new mxArray_tag();
  a->cls = cls;
  while (dims.size() < 2) dims.push_back(1);
  a->dims = dims;
  const size_t n = numel(a);
  if (cls == mxDOUBLE_CLASS) a->re.assign(n, 0.0);
  if (cls == mxLOGICAL_CLASS) a->lg.reset(new bool[n ? n : 1]());
  return a;
}
static std::vector<void (*)(void)> g_at_exit;
void mex_emul_run_at_exit() {
  for (auto f : g_at_exit) f();
  g_at_exit.clear();
}

extern "C" {
double mxGetScalar(const mxArray *a) {
  if (a->cls == mxDOUBLE_CLASS && !a->re.empty()) return a->re[0];
  if (a->cls == mxLOGICAL_CLASS && numel(a)) return a->lg[0] ? 1.0 : 0.0;
  if (a->cls == mxCHAR_CLASS && !a->chars.empty()) return (double)(unsigned char)a->chars[0];
  return 0.0;
}
double *mxGetPr(const mxArray *a) { return a->cls == mxDOUBLE_CLASS ? const_cast<double *>(a->re.data()) : nullptr; }
bool mxIsDouble(const mxArray *a) { return a->cls == mxDOUBLE_CLASS; }
bool mxIsComplex(const mxArray *) { return false; }
bool mxIsEmpty(const mxArray *a) { return numel(a) == 0; }
bool mxIsStruct(const mxArray *a) { return a->cls == mxSTRUCT_CLASS; }
bool mxIsChar(const mxArray *a) { return a->cls == mxCHAR_CLASS; }
static int field_no(const mxArray *a, const char *name) {
  for (size_t f = 0; f < a->fields.size(); ++f)
    if (a->fields[f] == name) return (int)f;
  return -1;
}
mxArray *mxGetFieldByNumber(const mxArray *a, mwIndex i, int f) {
  if (a->cls != mxSTRUCT_CLASS || f < 0 || (size_t)f >= a->fields.size() || i >= numel(a)) return nullptr;
  return a->vals[i * a->fields.size() + (size_t)f];
}
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name) { return mxGetFieldByNumber(a, i, field_no(a, name)); }
const char *mxGetFieldNameByNumber(const mxArray *a, int f) {
  return (a->cls == mxSTRUCT_CLASS && f >= 0 && (size_t)f < a->fields.size()) ? a->fields[(size_t)f].c_str() : nullptr;
}
int mxGetNumberOfFields(const mxArray *a) { return a->cls == mxSTRUCT_CLASS ? (int)a->fields.size() : 0; }
int mxAddField(mxArray *a, const char *name) {
  if (a->cls != mxSTRUCT_CLASS) return -1;
  const size_t n = numel(a), nf = a->fields.size();
  std::vector<mxArray *> v(n * (nf + 1), nullptr);
  for (size_t i = 0; i < n; ++i)
    for (size_t f = 0; f < nf; ++f) v[i * (nf + 1) + f] = a->vals[i * nf + f];
  a->vals.swap(v);
  a->fields.push_back(name);
  return (int)nf;
}
void mxSetField(mxArray *a, mwIndex i, const char *name, mxArray *v) {
  const int f = field_no(a, name);
  if (f < 0 || i >= numel(a)) return;
  a->vals[i * a->fields.size() + (size_t)f] = v;  // like MATLAB: the previous value is NOT freed
}
mxArray *mxDuplicateArray(const mxArray *a) {
  mxArray *b = new mxArray_tag();
  b->cls = a->cls; b->dims = a->dims; b->re = a->re; b->chars = a->chars; b->fields = a->fields;
  if (a->cls == mxLOGICAL_CLASS) {
    const size_t n = numel(a);
    b->lg.reset(new bool[n ? n : 1]());
    for (size_t i = 0; i < n; ++i) b->lg[i] = a->lg[i];
  }
  for (mxArray *v : a->vals) b->vals.push_back(v ? mxDuplicateArray(v) : nullptr);
  return b;
}
size_t mxGetNumberOfElements(const mxArray *a) { return numel(a); }
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) {   // product of dimensions 2..end, as documented
  size_t n = 1;
  for (size_t d = 1; d < a->dims.size(); ++d) n *= a->dims[d];
  return n;
}
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->dims.size(); }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims.data(); }
int mxGetString(const mxArray *a, char *buf, mwSize n) {
  if (a->cls != mxCHAR_CLASS || n == 0) return 1;
  const size_t k = a->chars.size() < n - 1 ? a->chars.size() : n - 1;
  memcpy(buf, a->chars.data(), k);
  buf[k] = 0;
  return a->chars.size() > n - 1;
}
double mxGetNaN(void) { return std::nan(""); }
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) { return make(mxDOUBLE_CLASS, {m, n}); }
mxArray *mxCreateNumericArray(mwSize nd, const mwSize *d, mxClassID cls, mxComplexity) {
  return make(cls, std::vector<mwSize>(d, d + nd));
}
mxArray *mxCreateLogicalMatrix(mwSize m, mwSize n) { return make(mxLOGICAL_CLASS, {m, n}); }
mxArray *mxCreateDoubleScalar(double v) {
  mxArray *a = make(mxDOUBLE_CLASS, {1, 1});
  a->re[0] = v;
  return a;
}
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nf, const char **names) {
  mxArray *a = make(mxSTRUCT_CLASS, {m, n});
  for (int f = 0; f < nf; ++f) a->fields.push_back(names[f]);
  a->vals.assign(numel(a) * (size_t)nf, nullptr);
  return a;
}
mxLogical *mxGetLogicals(const mxArray *a) { return a->cls == mxLOGICAL_CLASS ? a->lg.get() : nullptr; }
void mxDestroyArray(mxArray *a) {
  if (!a) return;
  for (mxArray *v : a->vals) mxDestroyArray(v);
  delete a;
}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw MexError{id ? id : "", buf};
}
int mexAtExit(void (*f)(void)) {
  g_at_exit.push_back(f);
  return 0;
}
}  // extern "C"

mxArray *mex_emul_char(const std::string &s) {
  mxArray *a = make(mxCHAR_CLASS, {1, s.size()});
  a->chars = s;
  return a;
}
