// Concrete mxArray of the MEX emulator (tests/mexhost/mex_emul.cpp).  Test infrastructure.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "mex.h"   // tests/stubs/mex.h: the API subset matlab/epi_mex.cpp uses

struct mxArray_tag {
  mxClassID cls = mxUNKNOWN_CLASS;
  std::vector<mwSize> dims;
  std::vector<double> re;            // mxDOUBLE_CLASS, column-major
  std::unique_ptr<bool[]> lg;        // mxLOGICAL_CLASS
  std::string chars;                 // mxCHAR_CLASS (1 x n)
  std::vector<std::string> fields;   // mxSTRUCT_CLASS
  std::vector<mxArray *> vals;       // [element][field]
};
struct MexError { std::string id, msg; };
void mex_emul_run_at_exit();
mxArray *mex_emul_char(const std::string &s);
