// stand-in for <cuda_runtime_api.h> in the syntax check of the jax.ffi handlers (only cudaStream_t is needed)
#pragma once
typedef struct CUstream_st* cudaStream_t;
