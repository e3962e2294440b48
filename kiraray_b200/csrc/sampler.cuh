// sampler.cuh -- PCG32 sampler, bit-exact with reference src/core/sampler.h:13-91.
// Only `state` is per pixel: inc = (sampleIndex << 1) | 1 with sampleIndex = frameIndex * spp
// (integrator.cpp:217) is the same for every pixel of a frame, so it is a kernel argument and the
// per-pixel RNG record shrinks from 16 to 8 bytes.
#pragma once
#include "krr_math.cuh"

namespace krr {

#define KRR_PCG32_MULT 0x5851f42d4c957f2dULL

// interleave_32bit, src/util/hash.h:12-27
KRR_HD uint32_t interleave32(uint32_t vx, uint32_t vy) {
	uint32_t x = vx & 0x0000ffff, y = vy & 0x0000ffff;
	x = (x | (x << 8)) & 0x00FF00FF; x = (x | (x << 4)) & 0x0F0F0F0F; x = (x | (x << 2)) & 0x33333333; x = (x | (x << 1)) & 0x55555555;
	y = (y | (y << 8)) & 0x00FF00FF; y = (y | (y << 4)) & 0x0F0F0F0F; y = (y | (y << 2)) & 0x33333333; y = (y | (y << 1)) & 0x55555555;
	return x | (y << 1);
}

struct Pcg {
	uint64_t state, inc;

	KRR_HD uint32_t nextUint() { // sampler.h:58-64
		uint64_t old = state;
		state		 = old * KRR_PCG32_MULT + inc;
		uint32_t xs	 = (uint32_t) (((old >> 18u) ^ old) >> 27u);
		uint32_t rot = (uint32_t) (old >> 59u);
		return (xs >> rot) | (xs << ((~rot + 1u) & 31));
	}
	KRR_HD void setSeed(uint64_t initstate, uint64_t initseq) { // sampler.h:22-28
		state = 0U;
		inc	  = (initseq << 1u) | 1u;
		nextUint();
		state += initstate;
		nextUint();
	}
	KRR_HD void setPixelSample(uint32_t px, uint32_t py, uint32_t sampleIndex) { setSeed(interleave32(px, py), sampleIndex); }
	KRR_HD void advance(int64_t delta) { // sampler.h:42-55
		uint64_t cur_mult = KRR_PCG32_MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		while (delta > 0) {
			if (delta & 1) {
				acc_mult *= cur_mult;
				acc_plus = acc_plus * cur_mult + cur_plus;
			}
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta /= 2;
		}
		state = acc_mult * state + acc_plus;
	}
	KRR_HD float get1D() { // nextFloat, sampler.h:77-86: [1,2) mantissa trick, exact subtraction
		uint32_t u = (nextUint() >> 9) | 0x3f800000u;
#ifdef __CUDA_ARCH__
		return __fsub_rn(__uint_as_float(u), 1.0f);
#else
		float f;
		memcpy(&f, &u, 4);
		return f - 1.0f;
#endif
	}
};

} // namespace krr
