// Host BVH build timing (no GPU needed): g++ -O3 -std=c++17 -pthread -Ibidirectional-pathtracing_b200/csrc scripts/bvh_build_bench.cpp bidirectional-pathtracing_b200/csrc/bvh8.cpp -o /tmp/bvh_bench && BDPT_BVH_THREADS=16 BDPT_BVH_TIMING=1 /tmp/bvh_bench 100000
#include "bvh8.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
using namespace bdpt;
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 100000;
  std::vector<BuildPrim> p(n);
  std::mt19937 g(1); std::uniform_real_distribution<float> u(0, 1);
  for (int i = 0; i < n; i++) { float c[3] = {u(g), u(g), u(g)}; for (int a = 0; a < 3; a++) { float e = 0.01f * u(g); p[i].lo[a] = c[a] - e; p[i].hi[a] = c[a] + e; } }
  for (int rep = 0; rep < 3; rep++) {
    Bvh8 b; auto t0 = std::chrono::steady_clock::now();
    build_bvh8(p, &b);
    printf("n=%d %.1f ms nodes %zu\n", n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), b.nodes.size() / 20);
  }
}
