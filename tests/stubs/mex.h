/* Minimal stand-in for MATLAB's mex.h: ONLY the declarations matlab/epi_mex.cpp uses, with the
 * documented MEX C API signatures.  It exists so the gateway can be type-checked in an image that
 * has neither MATLAB nor Octave (tests/test_capi_cpu.py); it is never linked or shipped. */
#ifndef EPI_TEST_STUB_MEX_H
#define EPI_TEST_STUB_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef bool mxLogical;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS,
               mxDOUBLE_CLASS, mxSINGLE_CLASS } mxClassID;
double mxGetScalar(const mxArray *);
double *mxGetPr(const mxArray *);
bool mxIsDouble(const mxArray *);
bool mxIsComplex(const mxArray *);
bool mxIsEmpty(const mxArray *);
bool mxIsStruct(const mxArray *);
bool mxIsChar(const mxArray *);
mxArray *mxGetField(const mxArray *, mwIndex, const char *);
mxArray *mxGetFieldByNumber(const mxArray *, mwIndex, int);
const char *mxGetFieldNameByNumber(const mxArray *, int);
int mxGetNumberOfFields(const mxArray *);
int mxAddField(mxArray *, const char *);
void mxSetField(mxArray *, mwIndex, const char *, mxArray *);
mxArray *mxDuplicateArray(const mxArray *);
size_t mxGetNumberOfElements(const mxArray *);
size_t mxGetM(const mxArray *);
size_t mxGetN(const mxArray *);
mwSize mxGetNumberOfDimensions(const mxArray *);
const mwSize *mxGetDimensions(const mxArray *);
int mxGetString(const mxArray *, char *, mwSize);
double mxGetNaN(void);
mxArray *mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray *mxCreateNumericArray(mwSize, const mwSize *, mxClassID, mxComplexity);
mxArray *mxCreateLogicalMatrix(mwSize, mwSize);
mxArray *mxCreateDoubleScalar(double);
mxArray *mxCreateStructMatrix(mwSize, mwSize, int, const char **);
mxLogical *mxGetLogicals(const mxArray *);
void mxDestroyArray(mxArray *);
void mexErrMsgIdAndTxt(const char *, const char *, ...);
int mexAtExit(void (*)(void));
/* the gateway entry point has C linkage in the real mex.h too */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
