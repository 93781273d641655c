// tma_host.cuh -- host helpers shared by the tensor-core kernels (defined in gemm.cu)
#pragma once
#include <cuda.h>
namespace nunerf {
// bf16 row-major matrix [rows, cols] with row pitch `ld` elements; box = box_cols x box_rows, 128B swizzle
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows);
int num_sms();
int env_int(const char* name, int dflt);
}  // namespace nunerf
