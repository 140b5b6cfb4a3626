// Micro-benchmark: issue rate of FADD / FADD2 / predicated-off FADD on sm_100a (per SM, lanes per clock).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, unsigned mask) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float v0 = out[0], v1 = out[1];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) { a[i] += v0; a[i + 1] += v1; }
      if (MODE == 1) {
        asm volatile("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
            : "+f"(a[i]), "+f"(a[i + 1]) : "f"(v0), "f"(v1));
      }
      if (MODE == 2) { if (mask & (1u << (i & 7))) { a[i] += v0; a[i + 1] += v1; } }           // predicated (mask = 0: all off)
      if (MODE == 3) {
        if (mask & (1u << (i & 7)))
          asm volatile("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
              : "+f"(a[i]), "+f"(a[i + 1]) : "f"(v0), "f"(v1));
      }
      if (MODE == 4) {   // one FADD2 + two FADD per 4 floats
        if ((i & 2) == 0)
          asm volatile("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
              : "+f"(a[i]), "+f"(a[i + 1]) : "f"(v0), "f"(v1));
        else { a[i] += v0; a[i + 1] += v1; }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 2 + 148 * 1024)[0] = t1 - t0;
}
int main() {
  float* d; cudaMalloc(&d, (2 + 148 * 1024 + 16) * 4); cudaMemset(d, 0, (2 + 148 * 1024 + 16) * 4);
  const int iters = 4096;
  const char* names[] = {"FADD scalar", "FADD2", "FADD predicated-off", "FADD2 predicated-off", "FADD2 + 2 FADD", "FADD predicated half"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 1024>>>(d, iters, 0xff);
      if (mode == 1) k<1><<<148, 1024>>>(d, iters, 0xff);
      if (mode == 2) k<2><<<148, 1024>>>(d, iters, 0);
      if (mode == 3) k<3><<<148, 1024>>>(d, iters, 0);
      if (mode == 4) k<4><<<148, 1024>>>(d, iters, 0xff);
      if (mode == 5) k<2><<<148, 1024>>>(d, iters, 0x55);
      cudaDeviceSynchronize();
    }
    long long cyc; cudaMemcpy(&cyc, d + 2 + 148 * 1024, 8, cudaMemcpyDeviceToHost);
    // per SM: 32 warps * iters * 16 float adds (slots) per lane
    double slots = 32.0 * 32 * iters * 16;
    printf("%-24s %lld cycles  -> %.1f float-add slots / clk / SM\n", names[mode], cyc, slots / cyc);
  }
  return 0;
}
