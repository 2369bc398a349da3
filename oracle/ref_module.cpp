// TEST INFRASTRUCTURE ONLY -- not part of the product path.
// Five-line pybind11 module stub that exposes the UNMODIFIED reference tokenizer
// (compiled from /root/reference/src/{tokenize.cpp,omp.cpp,fxstats.cpp} where they lie; see
// oracle/Makefile) under the module name `ref_cbioseq`, so it can be imported next
// to the product's own `cbioseq` extension.  Mirrors what src/bioseq.cpp:6-11 does
// minus poa (which needs spoa and is not on the hot path).  fxstats (FlatFile, the on-disk
// packed-sequence format that feeds the path; needs only zlib) is included.
#include "bioseq.h"
void init_omp_helpers(py::module &m);
void init_fxstats(py::module &m);
PYBIND11_MODULE(ref_cbioseq, m) {
    init_tokenize(m);
    init_omp_helpers(m);
    init_fxstats(m);
}
