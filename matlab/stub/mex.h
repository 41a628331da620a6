/* STUB of MATLAB's mex.h / matrix.h -- compile check only.
 * Neither MATLAB nor GNU Octave exists in the build image, so chest_mex.c is compiled against
 * these declarations (the documented MATLAB C Matrix API, split-complex "separate" layout that
 * R2013b/R2016a and Octave's mkoctfile --mex provide).  Nothing here is linked or executed. */
#ifndef CHEST_STUB_MEX_H
#define CHEST_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxUINT8_CLASS = 9, mxINT32_CLASS = 12, mxUINT32_CLASS = 13, mxUINT64_CLASS = 15 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
double* mxGetPr(const mxArray*);
double* mxGetPi(const mxArray*);
void* mxGetData(const mxArray*);
mwIndex* mxGetIr(const mxArray*);
mwIndex* mxGetJc(const mxArray*);
mwSize mxGetM(const mxArray*);
mwSize mxGetN(const mxArray*);
mwSize mxGetNzmax(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsSparse(const mxArray*);
int mxIsDouble(const mxArray*);
int mxIsChar(const mxArray*);
int mxIsEmpty(const mxArray*);
double mxGetScalar(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateSparse(mwSize, mwSize, mwSize, mxComplexity);
mxArray* mxCreateCellMatrix(mwSize, mwSize);
void mxSetCell(mxArray*, mwIndex, mxArray*);
void* mxMalloc(size_t);
void* mxCalloc(size_t, size_t);
void mxFree(void*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexLock(void);
void mexUnlock(void);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
