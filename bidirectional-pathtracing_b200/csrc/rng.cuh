// rng.cuh — counter-keyed random streams.  One PCG-RXS-M-XS 32/32 stream (32-bit LCG state, its increment and start
// taken from a 64-bit hash of the key: 7 integer instructions a draw where the 64/32 generator of round 1 took ~16 —
// 64-bit multiplies are four IMADs on the GPU — walk -2.3 %; a stream here is at most a few dozen draws) per
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
  uint32_t state, inc;
  __host__ __device__ inline void begin(uint64_t seed, uint64_t pixel, uint64_t sample, uint32_t stream) {
    uint64_t h = mix64(seed ^ mix64(pixel * 0x9E3779B97F4A7C15ull + sample) ^
                       (0xD1B54A32D192ED03ull * (uint64_t)(stream + 1)));
    state = (uint32_t)h; inc = (uint32_t)(h >> 32) | 1u;   // (an odd increment: full period 2^32)
  }
  __host__ __device__ inline uint32_t next() {
    uint32_t old = state;
    state = old * 747796405u + inc;
    uint32_t word = ((old >> ((old >> 28u) + 4u)) ^ old) * 277803737u;
    return (word >> 22u) ^ word;
  }
  __host__ __device__ inline float uniform() { return ((float)(next() >> 9) + 0.5f) * (1.0f / 8388608.0f); }
  __host__ __device__ inline int pick(int n) {
    int k = (int)(uniform() * (float)n);
    return k < n ? k : n - 1;
  }
};

}  // namespace bdpt
