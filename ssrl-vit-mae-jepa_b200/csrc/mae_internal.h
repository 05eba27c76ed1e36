// Internal declarations shared by the .cu translation units of libmae_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mae_b200.h"

namespace mae {
int make_tmap_bf16(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows);
}

#define MAE_CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? MAE_OK : MAE_ERR_LAUNCH)
