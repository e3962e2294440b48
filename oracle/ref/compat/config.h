/* Stand-in for the reference's CMake-generated config.h (template: src/core/config.in.h). */
#pragma once
#define KRR_PROJECT_NAME "krr-oracle"
#define KRR_PROJECT_DIR "/root/reference"
#define KRR_BUILD_TYPE "Release"
#define KRR_VULKAN_ROOT ""
#define KRR_PYTORCH_ROOT ""
#define KRR_BUILD_STARLIGHT 0
#define KRR_RENDER_SPECTRAL 1
#define KRR_ENABLE_PYTORCH 0
#define KRR_ENABLE_PROFILE 0
#define KRR_N_SPECTRUM_SAMPLES 4
#define KRR_CLIPSPACE_RIGHTHANDED 1
#define KRR_CLIPSPACE_Z_FROM_ZERO 1
#define KRR_DEFAULT_RND_SEED 7272
#define OPTIX_LOG_SIZE 4096
