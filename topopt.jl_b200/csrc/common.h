// Shared declarations of libtopopt_cuda (host side).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/topopt_cuda.h"

namespace topopt {

struct GridDims {
  int dim;
  int64_t nx, ny, nz;  // elements per axis (nz = 1 in 2-D)
  int64_t NX, NY, NZ;  // nodes per axis (NZ = 1 in 2-D)
  int64_t nnodes, nel;
};

extern thread_local std::string g_last_error;
void set_error(const std::string& msg);
bool check_dims(int32_t dim, int32_t ncomp, const int64_t* nels);
GridDims make_dims(int32_t dim, const int64_t* nels);
void cell_nodes(const GridDims& g, int64_t e, int64_t* out);
std::vector<int64_t> ferrite_node_blocks(const GridDims& g);
int64_t pattern_nnz(const GridDims& g, int ncomp);

}  // namespace topopt
