// development probe: glm::reduce192 vs gl::reduce_hint on the device
#include <cstdio>
#include "../../gnark-plonky2-verifier_b200/csrc/poseidon_gl_macro.cuh"
using namespace gpw;
__global__ void k(unsigned long long seed, unsigned long long* bad, unsigned long long* ex) {
  unsigned long long s = seed + (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
  auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int it = 0; it < 1000; it++) {
    glm::U192 v{{next(), next(), next() % gl::P}};
    if (it % 3 == 0) v.l[2] = 0;
    uint64_t x[4] = {v.l[0], v.l[1], v.l[2], 0}, q[4], r, q0, q1, r2;
    gl::reduce_hint(x, q, r);
    glm::reduce192(v, q0, q1, r2);
    if (r != r2 || q0 != q[0] || q1 != q[1]) {
      if (atomicAdd(bad, 1ull) == 0) { ex[0] = v.l[0]; ex[1] = v.l[1]; ex[2] = v.l[2]; ex[3] = q[0]; ex[4] = q[1]; ex[5] = r; ex[6] = q0; ex[7] = q1; ex[8] = r2; }
    }
  }
}
int main() {
  unsigned long long *bad, *ex;
  cudaMallocManaged(&bad, 8); cudaMallocManaged(&ex, 80);
  *bad = 0;
  k<<<64, 128>>>(12345, bad, ex);
  cudaDeviceSynchronize();
  printf("bad=%llu\n", *bad);
  if (*bad) printf("x=%016llx %016llx %016llx\n ref q=%016llx %016llx r=%016llx\n got q=%016llx %016llx r=%016llx\n", ex[0], ex[1], ex[2], ex[3], ex[4], ex[5], ex[6], ex[7], ex[8]);
  return 0;
}
