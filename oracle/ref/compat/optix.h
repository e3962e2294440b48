/* Minimal stand-in for <optix.h>: only the handle types the reference's headers mention. */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
typedef unsigned long long OptixTraversableHandle;
typedef struct OptixDeviceContext_t *OptixDeviceContext;
#define OPTIX_VERSION 80000
