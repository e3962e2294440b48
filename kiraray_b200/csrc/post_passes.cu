// post_passes.cu -- device kernels of the render passes that follow the path tracer on the reference's
// example configs: AccumulatePass in double precision, ErrorMeasurePass metrics and
// ToneMappingPass operators.  All of them stream the RGBA32F film once (HBM-bound, 16 B/pixel read and,
// where the pass rewrites the film, 16 B/pixel written); grids are a multiple of the SM count.
//
//   k_accumulate_f64   acculumate<double>           src/render/passes/accumulate/accumulate.cu:30-52
//   k_read_average     AccumulatePass::saveImage    accumulate.cu:90-111
//   k_error_metric     calc_metric                  src/render/passes/errormeasure/metrics.cu:64-128
//   k_tonemap          ToneMappingPass::render      src/render/passes/tonemapping/tonemapping.cu:11-85
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "krr_wfpt.h"

namespace krr {
void setLastError(const char *fmt, ...); // api.cu
namespace {

constexpr int kPostBlock = 256;
// These entry points take no handle: they work on the CURRENT device of the calling thread (the device the
// film pointer belongs to), so nothing device-specific may be cached process-wide.
int postGrid() {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return sms * 8;
}

__global__ void k_accumulate_f64(double4 *accum, float4 *film, long long n, unsigned long long accumCount, unsigned long long maxAccum, int movingAverage) {
	for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		const double w = 1.0 / (double) (accumCount + 1);
		const float4 cf = film[i];
		const double4 c = make_double4(cf.x, cf.y, cf.z, cf.w);
		double4 a;
		if (accumCount > 0) {
			a = accum[i];
			if (movingAverage) a = make_double4(a.x * (1 - w) + c.x * w, a.y * (1 - w) + c.y * w, a.z * (1 - w) + c.z * w, a.w * (1 - w) + c.w * w);
			else if (!maxAccum || accumCount < maxAccum) a = make_double4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
		} else a = c;
		accum[i] = a;
		film[i]	 = movingAverage ? make_float4((float) a.x, (float) a.y, (float) a.z, (float) a.w)
								 : make_float4((float) (a.x * w), (float) (a.y * w), (float) (a.z * w), (float) (a.w * w));
	}
}

// accumulated sum * (1 / count) as float (saveImage)
__global__ void k_read_average(const void *accum, int isDouble, float4 *out, long long n, float weight) {
	for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		if (isDouble) {
			const double4 a = ((const double4 *) accum)[i];
			out[i] = make_float4((float) (a.x * weight), (float) (a.y * weight), (float) (a.z * weight), (float) (a.w * weight));
		} else {
			const float4 a = ((const float4 *) accum)[i];
			out[i] = make_float4(a.x * weight, a.y * weight, a.z * weight, a.w * weight);
		}
	}
}

// ---- ErrorMeasurePass: per-pixel error (metrics.cu:64-97), clamped at 100 (:20, 117-119), then summed.
// The reference reduces the per-pixel floats with thrust::reduce in float; here each thread adds its
// pixels in double and the block totals are added with one atomicAdd(double) per block, so the result
// does not depend on a second pass over an intermediate buffer.
__device__ bool invalid(float r, float g, float b) { return isinf(r) || isinf(g) || isinf(b) || isnan(r) || isnan(g) || isnan(b); }
__device__ float pixelError(float4 yc, float4 rc, int metric) {
	const float y[3] = {yc.x, yc.y, yc.z}, ref[3] = {rc.x, rc.y, rc.z};
	if (invalid(ref[0], ref[1], ref[2])) return 0.f; // CHECK_INVALID(ref)
	float e[3];
	for (int c = 0; c < 3; c++) {
		const float d = fabsf(y[c] - ref[c]);
		switch (metric) {
			case KRR_METRIC_MSE: e[c] = d * d; break;					  // (y - ref).abs().pow(2)
			case KRR_METRIC_MAPE: e[c] = d / (ref[c] + 0.f); break;	  // ERROR_EPS = 0
			case KRR_METRIC_SMAPE: e[c] = d / (ref[c] + y[c] + 0.f); break;
			default: e[c] = ref[c] == 0.f ? 0.f : (d / ref[c]) * (d / ref[c]); // rel_mse with ERROR_EPS == 0
		}
	}
	const float err = (e[0] + e[1] + e[2]) / 3.f; // Array3f::mean()
	return fminf(err, 100.f);					  // CLAMP_PIXEL_ERROR_THRESHOLD (fminf drops a NaN error like min() on the device)
}
__global__ void __launch_bounds__(kPostBlock) k_error_metric(const float4 *film, const float4 *reference, long long n, int metric, double *sum) {
	__shared__ double part[kPostBlock / 32];
	double acc = 0;
	for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
		acc += (double) pixelError(film[i], reference[i], metric);
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double t = 0;
		for (int w = 0; w < kPostBlock / 32; w++) t += part[w];
		atomicAdd(sum, t);
	}
}

// ---- ToneMappingPass operators (tonemapping.cu:11-46) ----
__device__ float3 toneMapAces(float3 c) {
	c = make_float3(c.x * 0.6f, c.y * 0.6f, c.z * 0.6f);
	const float A = 2.51f, B = 0.03f, C = 2.43f, D = 0.59f, E = 0.14f;
	auto f = [&](float x) { return fminf(fmaxf((x * (A * x + B)) / (x * (C * x + D) + E), 0.f), 1.f); };
	return make_float3(f(c.x), f(c.y), f(c.z));
}
__device__ float3 toneMapReinhard(float3 c) {
	const float lum = c.x * 0.299f + c.y * 0.587f + c.z * 0.114f, reinhard = lum / (lum + 1);
	auto f = [&](float x) { return lum == 0.f ? 0.f : (x * reinhard) / lum; }; // safeDiv
	return make_float3(f(c.x), f(c.y), f(c.z));
}
__device__ float3 toneMapUC2(float3 c) {
	const float A = 0.22f, B = 0.3f, C = 0.1f, D = 0.2f, E = 0.01f, F = 0.3f;
	auto f = [&](float x) { return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - (E / F); };
	return make_float3(f(c.x), f(c.y), f(c.z));
}
__device__ float3 toneMapHejiHableAlu(float3 c) {
	auto f = [&](float x) {
		x = fmaxf(x - 0.004f, 0.f);
		x = (x * (6.2f * x + 0.5f)) / (x * (6.2f * x + 1.7f) + 0.06f);
		return powf(x, 2.2f); // "Result includes sRGB conversion"
	};
	return make_float3(f(c.x), f(c.y), f(c.z));
}
__global__ void k_tonemap(float4 *film, long long n, int op, float exposure, int useGamma) {
	for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		const float4 p = film[i];
		float3 c = make_float3(p.x * exposure, p.y * exposure, p.z * exposure);
		switch (op) {
			case KRR_TONEMAP_REINHARD: c = toneMapReinhard(c); break;
			case KRR_TONEMAP_ACES: c = toneMapAces(c); break;
			case KRR_TONEMAP_UNCHARTED2: c = toneMapUC2(c); break;
			case KRR_TONEMAP_HEJIHABLE: c = toneMapHejiHableAlu(c); break;
			default: break;
		}
		if (useGamma) c = make_float3(powf(c.x, 0.45454545f), powf(c.y, 0.45454545f), powf(c.z, 0.45454545f));
		film[i] = make_float4(c.x, c.y, c.z, 1.f);
	}
}

int fail(int code, const char *msg) {
	setLastError("%s", msg);
	return code;
}
#define POST_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { setLastError("%s: %s", #x, cudaGetErrorString(e_)); return KRR_E_CUDA; } } while (0)

} // namespace
} // namespace krr

using namespace krr;

extern "C" int krr_accumulate_f64(double *accum, float *film, int64_t n, uint64_t accumCount, uint64_t maxAccum, int32_t movingAverage, void *stream) {
	if (!accum || !film || n <= 0) return fail(KRR_E_INVALID, "krr_accumulate_f64: bad argument");
	k_accumulate_f64<<<postGrid(), kPostBlock, 0, (cudaStream_t) stream>>>((double4 *) accum, (float4 *) film, n, accumCount, maxAccum, movingAverage);
	POST_OK(cudaGetLastError());
	return KRR_OK;
}

extern "C" int krr_accumulate_read_average(const void *accum, int32_t isDouble, uint64_t accumCount, int64_t n, float *out_host, void *stream) {
	if (!accum || !out_host || n <= 0 || accumCount == 0) return fail(KRR_E_INVALID, "krr_accumulate_read_average: bad argument (nothing accumulated?)");
	float4 *tmp = nullptr;
	POST_OK(cudaMalloc((void **) &tmp, (size_t) n * 16));
	k_read_average<<<postGrid(), kPostBlock, 0, (cudaStream_t) stream>>>(accum, isDouble, tmp, n, 1.0f / (float) accumCount);
	cudaError_t e = cudaMemcpyAsync(out_host, tmp, (size_t) n * 16, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
	cudaFree(tmp);
	if (e != cudaSuccess) { setLastError("krr_accumulate_read_average: %s", cudaGetErrorString(e)); return KRR_E_CUDA; }
	return KRR_OK;
}

extern "C" int krr_error_metric_f32(const float *film, const float *reference, int64_t n, int32_t metric, double *result_host, void *stream) {
	if (!film || !reference || !result_host || n <= 0) return fail(KRR_E_INVALID, "krr_error_metric_f32: bad argument");
	if (metric < 0 || metric > KRR_METRIC_REL_MSE) return fail(KRR_E_INVALID, "krr_error_metric_f32: unknown metric");
	// reduction scratch on the current device, freed before returning (the call synchronises anyway)
	double *dSum = nullptr;
	POST_OK(cudaMallocAsync((void **) &dSum, 8, (cudaStream_t) stream));
	double sum = 0;
	cudaError_t e = cudaMemsetAsync(dSum, 0, 8, (cudaStream_t) stream);
	if (e == cudaSuccess) {
		k_error_metric<<<postGrid(), kPostBlock, 0, (cudaStream_t) stream>>>((const float4 *) film, (const float4 *) reference, n, metric, dSum);
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) e = cudaMemcpyAsync(&sum, dSum, 8, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
	cudaFreeAsync(dSum, (cudaStream_t) stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
	if (e != cudaSuccess) { setLastError("krr_error_metric_f32: %s", cudaGetErrorString(e)); return KRR_E_CUDA; }
	*result_host = sum / (double) n;
	return KRR_OK;
}

extern "C" int krr_tonemap_f32(float *film, int64_t n, int32_t op, float exposure, int32_t useGamma, void *stream) {
	if (!film || n <= 0) return fail(KRR_E_INVALID, "krr_tonemap_f32: bad argument");
	if (op < 0 || op > KRR_TONEMAP_HEJIHABLE) return fail(KRR_E_INVALID, "krr_tonemap_f32: unknown operator");
	k_tonemap<<<postGrid(), kPostBlock, 0, (cudaStream_t) stream>>>((float4 *) film, n, op, exposure, useGamma);
	POST_OK(cudaGetLastError());
	return KRR_OK;
}
