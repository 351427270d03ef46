// michigan_b200 — internal helpers shared by the translation units of libmichigan_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/michigan_b200.h"

namespace mg {

// printf-style; stores the message in a thread-local buffer and returns `code`.
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
int encode_tensor_map(CUtensorMap* map, void* gaddr, CUtensorMapDataType dtype, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* estrides, CUtensorMapSwizzle swizzle);

inline int check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // namespace mg
