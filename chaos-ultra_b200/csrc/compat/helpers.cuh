/* src/main/cuda/helpers.cuh for a reference-style module: see chaos_compat_pre.cuh */
#include "chaos_compat_pre.cuh"
