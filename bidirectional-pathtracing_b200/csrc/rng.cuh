// rng.cuh — counter-keyed random streams.  One PCG-XSH-RR 64/32 stream per
// (seed, pixel, sample index, stream id); the fp64 oracle (oracle/bdpt_oracle.cpp, Rng,
// rng_mode 1) implements the identical generator so both sides draw the same numbers.
//
// The reference draws from process-global mt19937 engines (src/util/random_util.h:10-22),
// which makes it order- and thread-dependent; keying on (pixel, sample)
// instead makes the image independent of batching, of the GPU count and of scheduling.
//
// stream ids: 0 = eye subpath (pixel jitter first), 1 = light subpath,
//             2 + t = the fresh light sample of strategy (s = 1, t).
#pragma once
#include <stdint.h>

namespace bdpt {

__host__ __device__ inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

struct Rng {
  uint64_t state, inc;

  __host__ __device__ inline void begin(uint64_t seed, uint64_t pixel, uint64_t sample, uint32_t stream) {
    uint64_t h = mix64(seed ^ mix64(pixel * 0x9E3779B97F4A7C15ull + sample) ^
                       (0xD1B54A32D192ED03ull * (uint64_t)(stream + 1)));
    inc = (mix64(h) << 1) | 1ull;
    state = h;
  }
  __host__ __device__ inline uint32_t next() {
    uint64_t old = state;
    state = old * 6364136223846793005ull + inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
  // (k + 0.5) * 2^-23, k = top 23 bits: strictly inside (0,1) and exactly representable in fp32,
  // so the fp64 oracle and the fp32 kernels see the same value.
  __host__ __device__ inline float uniform() { return ((float)(next() >> 9) + 0.5f) * (1.0f / 8388608.0f); }
  __host__ __device__ inline int pick(int n) {
    int k = (int)(uniform() * (float)n);
    return k < n ? k : n - 1;
  }
};

}  // namespace bdpt
