// Record of one Muon-labelled matrix (built by ParamStore.muon_table, params.py) and the tensor-core Newton-Schulz entry.
#pragma once
#include <cuda_runtime.h>

namespace mpx {

struct MuonMat {
  long long off;       // offset of the parameter in the flat buffers
  int rows, cols;      // parameter shape (in, out)
  int r, c;            // oriented shape, r <= c
  int transposed;      // 1 if the oriented matrix is the transpose of the parameter
  int _pad;
  long long xoff;      // offset of this matrix in the X / X' workspaces (r*c floats)
  long long aoff;      // offset in the A / B workspaces (r*r floats)
};

// muon_ns_tc.cu: A = X X^T, B = c1 A + c2 A A, X' = c0 X + B X for every matrix of the table (three launches)
int muon_ns_tc_iteration(const MuonMat* mats, int nmat, int max_r, int max_c, const float* xin, float* xout, float* Abuf,
                         float* Bbuf, const float* norm2, int first, float c0, float c1, float c2, float eps, cudaStream_t s);

}  // namespace mpx
