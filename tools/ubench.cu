// ubench.cu -- B200 micro-benchmarks that drive the tile-kernel design:
//   (1) raw pipe rates: FFMA, FADD, FFMA2, FADD2, FMNMX, FSETP, MUFU, F2I
//   (2) candidate inner loops of the pair kernel (4 i-atoms in registers x
//       1024 j-atoms broadcast from shared memory), scalar and f32x2-packed
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/_build/ubench tools/ubench.cu
// Run on the GPU box; prints one line per test.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---------------------------------------------------------------- pipe rates
enum { OP_FFMA, OP_FADD, OP_FFMA2, OP_FADD2, OP_FMNMX, OP_FFMA_FMNMX, OP_FSETP,
       OP_MUFU, OP_F2I, OP_FFMA2_FMNMX, OP_IADD, OP_FFMA_IADD, OP_COUNT };
static const char *op_names[] = {"FFMA", "FADD", "FFMA2(x2 lanes)", "FADD2(x2 lanes)",
  "FMNMX", "FFMA+FMNMX 1:1", "FSETP+SEL", "MUFU.RSQ", "F2I", "FFMA2+FMNMX 1:1", "IADD3", "FFMA+IADD 1:1"};

template <int OP>
__global__ __launch_bounds__(256) void rate_kernel(float *out, int iters, float a, float b) {
  float v[8];
  uint64_t w[8];
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x + i; w[i] = pk(v[i], v[i] + 0.5f); q[i] = threadIdx.x + i; }
  const uint64_t a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == OP_FFMA) v[i] = __fmaf_rn(v[i], a, b);
        if (OP == OP_FADD) v[i] = __fadd_rn(v[i], a);
        if (OP == OP_FFMA2) w[i] = fma2(w[i], a2, b2);
        if (OP == OP_FADD2) w[i] = add2(w[i], a2);
        if (OP == OP_FMNMX) v[i] = fminf(v[i] , a) , a += 0.f;
        if (OP == OP_FFMA_FMNMX) { v[i] = __fmaf_rn(v[i], a, b); v[(i + 4) & 7] = fmaxf(v[(i + 4) & 7], b); }
        if (OP == OP_FSETP) q[i] += (v[i] < a + q[i]) ? 1 : 2;
        if (OP == OP_MUFU) v[i] = __frsqrt_rn(v[i]);
        if (OP == OP_F2I) q[i] += __float2int_rd(v[i] + q[i]);
        if (OP == OP_FFMA2_FMNMX) { w[i] = fma2(w[i], a2, b2); v[i] = fmaxf(v[i], b) + 0.f; }
        if (OP == OP_IADD) q[i] = q[i] * 3 + it;
        if (OP == OP_FFMA_IADD) { v[i] = __fmaf_rn(v[i], a, b); q[i] = (q[i] ^ it) + u; }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float lo, hi; unpk(w[i], lo, hi); s += v[i] + lo + hi + q[i]; }
  if (s == 123.456f) out[0] = s;
}

// ------------------------------------------------------- pair-loop variants
constexpr int TJ = 1024;
constexpr int RI = 4;
enum { V_GENERIC, V_WRAPPED, V_TWOIMG, V_NOPBC, V_GENERIC_X2, V_TWOIMG_X2, V_NOPBC_X2, V_TRI, V_COUNT };
static const char *v_names[] = {"generic ortho (magic rint) 15op", "wrapped L/2-|q| 12op",
  "two-image min 6add+3mnmx+3", "no-pbc/shifted 6op", "generic ortho f32x2", "two-image f32x2",
  "no-pbc f32x2", "triclinic brick 18op"};
static const double v_flops[] = {20, 20, 20, 8, 20, 20, 8, 26};

template <int V>
__global__ __launch_bounds__(256, 2) void pair_kernel(const float *__restrict__ g, int iters, float L, float cut, unsigned long long *out) {
  __shared__ __align__(16) float xs[TJ], ys[TJ], zs[TJ];
  for (int t = threadIdx.x; t < TJ; t += blockDim.x) {
    xs[t] = g[3 * t]; ys[t] = g[3 * t + 1]; zs[t] = g[3 * t + 2];
  }
  __syncthreads();
  float xi[RI], yi[RI], zi[RI], xb[RI], yb[RI], zb[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    const int a = (blockIdx.x * 256 + threadIdx.x + r * 977) % TJ;
    xi[r] = g[3 * a]; yi[r] = g[3 * a + 1]; zi[r] = g[3 * a + 2];
    xb[r] = xi[r] + (xi[r] < 0.5f * L ? L : -L);
    yb[r] = yi[r] + (yi[r] < 0.5f * L ? L : -L);
    zb[r] = zi[r] + (zi[r] < 0.5f * L ? L : -L);
  }
  const float invL = 1.0f / L, M = 12582912.0f, hL = 0.5f * L;
  const uint64_t invL2 = pk(invL, invL), M2 = pk(M, M), nM2 = pk(-M, -M), nL2 = pk(-L, -L);
  const float cxx = 0.5f * L, cyy = 0.3f * L;
  unsigned int cnt = 0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int j = 0; j < TJ; j += 4) {
      const float4 X = *reinterpret_cast<const float4 *>(xs + j);
      const float4 Y = *reinterpret_cast<const float4 *>(ys + j);
      const float4 Z = *reinterpret_cast<const float4 *>(zs + j);
      if (V == V_GENERIC || V == V_WRAPPED || V == V_TWOIMG || V == V_NOPBC || V == V_TRI) {
        const float xj[4] = {X.x, X.y, X.z, X.w}, yj[4] = {Y.x, Y.y, Y.z, Y.w}, zj[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float m = 1e30f;
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            float dx = __fsub_rn(xj[u], xi[r]), dy = __fsub_rn(yj[u], yi[r]), dz = __fsub_rn(zj[u], zi[r]);
            if (V == V_GENERIC) {
              const float nx = __fadd_rn(__fmaf_rn(dx, invL, M), -M), ny = __fadd_rn(__fmaf_rn(dy, invL, M), -M), nz = __fadd_rn(__fmaf_rn(dz, invL, M), -M);
              dx = __fmaf_rn(-L, nx, dx); dy = __fmaf_rn(-L, ny, dy); dz = __fmaf_rn(-L, nz, dz);
            } else if (V == V_WRAPPED) {
              dx = __fsub_rn(hL, fabsf(__fsub_rn(fabsf(dx), hL)));
              dy = __fsub_rn(hL, fabsf(__fsub_rn(fabsf(dy), hL)));
              dz = __fsub_rn(hL, fabsf(__fsub_rn(fabsf(dz), hL)));
            } else if (V == V_TWOIMG) {
              dx = fminf(fabsf(dx), fabsf(__fsub_rn(xj[u], xb[r])));
              dy = fminf(fabsf(dy), fabsf(__fsub_rn(yj[u], yb[r])));
              dz = fminf(fabsf(dz), fabsf(__fsub_rn(zj[u], zb[r])));
            } else if (V == V_TRI) {
              const float nz = __fadd_rn(__fmaf_rn(dz, invL, M), -M);
              dz = __fmaf_rn(-L, nz, dz); dy = __fmaf_rn(-cyy, nz, dy); dx = __fmaf_rn(-cxx, nz, dx);
              const float ny = __fadd_rn(__fmaf_rn(dy, invL, M), -M);
              dy = __fmaf_rn(-L, ny, dy); dx = __fmaf_rn(-cxx, ny, dx);
              const float nx = __fadd_rn(__fmaf_rn(dx, invL, M), -M);
              dx = __fmaf_rn(-L, nx, dx);
            }
            const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            m = fminf(m, r2);
          }
          if (m < cut) ++cnt;
        }
      } else {
        // packed over j pairs: (j0,j1) and (j2,j3)
        const uint64_t xj2[2] = {pk(X.x, X.y), pk(X.z, X.w)}, yj2[2] = {pk(Y.x, Y.y), pk(Y.z, Y.w)}, zj2[2] = {pk(Z.x, Z.y), pk(Z.z, Z.w)};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float m = 1e30f;
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            const uint64_t xi2 = pk(xi[r], xi[r]), yi2 = pk(yi[r], yi[r]), zi2 = pk(zi[r], zi[r]);
            uint64_t dx = sub2(xj2[u], xi2), dy = sub2(yj2[u], yi2), dz = sub2(zj2[u], zi2);
            if (V == V_GENERIC_X2) {
              const uint64_t nx = add2(fma2(dx, invL2, M2), nM2), ny = add2(fma2(dy, invL2, M2), nM2), nz = add2(fma2(dz, invL2, M2), nM2);
              dx = fma2(nL2, nx, dx); dy = fma2(nL2, ny, dy); dz = fma2(nL2, nz, dz);
            } else if (V == V_TWOIMG_X2) {
              const uint64_t xb2 = pk(xb[r], xb[r]), yb2 = pk(yb[r], yb[r]), zb2 = pk(zb[r], zb[r]);
              const uint64_t ex = sub2(xj2[u], xb2), ey = sub2(yj2[u], yb2), ez = sub2(zj2[u], zb2);
              float a0, a1, b0, b1;
              unpk(dx, a0, a1); unpk(ex, b0, b1); dx = pk(fminf(fabsf(a0), fabsf(b0)), fminf(fabsf(a1), fabsf(b1)));
              unpk(dy, a0, a1); unpk(ey, b0, b1); dy = pk(fminf(fabsf(a0), fabsf(b0)), fminf(fabsf(a1), fabsf(b1)));
              unpk(dz, a0, a1); unpk(ez, b0, b1); dz = pk(fminf(fabsf(a0), fabsf(b0)), fminf(fabsf(a1), fabsf(b1)));
            }
            const uint64_t r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float lo, hi;
            unpk(r2, lo, hi);
            m = fminf(m, fminf(lo, hi));
          }
          if (m < cut) ++cnt;
        }
      }
    }
  }
  if (cnt == 0xffffffffu) out[0] = cnt;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = cnt;
}

template <int OP>
static void run_rate(int sms, float *dout) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = sms * 8, iters = 2048;
  rate_kernel<OP><<<grid, 256>>>(dout, 16, 1.0001f, 0.5f);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    rate_kernel<OP><<<grid, 256>>>(dout, iters, 1.0001f, 0.5f);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double warp_instr = (double)iters * 64 * (256 / 32) * grid;  // per op kind
  const double per_sm_clk = warp_instr / sms / (best * 1e-3);         // warp-instr / s / SM
  printf("RATE %-18s %8.3f ms  %7.2f Gwarp-instr/s/SM  (=%.2f warp-instr/clk/SM @1.965GHz, %.2f @1.5GHz)\n",
         op_names[OP], best, per_sm_clk / 1e9, per_sm_clk / 1.965e9, per_sm_clk / 1.5e9);
}

template <int V>
static void run_pair(int sms, const float *g, unsigned long long *dout) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = sms * 2, iters = 64;
  pair_kernel<V><<<grid, 256>>>(g, 2, 40.f, 225.f, dout);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    pair_kernel<V><<<grid, 256>>>(g, iters, 40.f, 225.f, dout);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double pairs = (double)iters * TJ * RI * 256.0 * grid;
  const double pps = pairs / (best * 1e-3);
  printf("PAIR %-34s %8.3f ms  %.3e pairs/s  %.2f pairs/clk/SM@1.965  %.1f TFLOP/s-equiv(%g flop/pair)\n",
         v_names[V], best, pps, pps / sms / 1.965e9, pps * v_flops[V] / 1e12, v_flops[V]);
}

int main() {
  int sms = 0, dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  printf("SMs %d  clockRate %d kHz\n", sms, clk);
  float *dout; CK(cudaMalloc(&dout, 64));
  unsigned long long *dout2; CK(cudaMalloc(&dout2, 64));
  float *h = (float *)malloc(3 * TJ * sizeof(float));
  srand(1);
  for (int i = 0; i < 3 * TJ; ++i) h[i] = 40.f * (rand() / (float)RAND_MAX);
  float *g; CK(cudaMalloc(&g, 3 * TJ * sizeof(float)));
  CK(cudaMemcpy(g, h, 3 * TJ * sizeof(float), cudaMemcpyHostToDevice));
  run_rate<OP_FFMA>(sms, dout); run_rate<OP_FADD>(sms, dout); run_rate<OP_FFMA2>(sms, dout);
  run_rate<OP_FADD2>(sms, dout); run_rate<OP_FMNMX>(sms, dout); run_rate<OP_FFMA_FMNMX>(sms, dout);
  run_rate<OP_FSETP>(sms, dout); run_rate<OP_MUFU>(sms, dout); run_rate<OP_F2I>(sms, dout);
  run_rate<OP_FFMA2_FMNMX>(sms, dout); run_rate<OP_IADD>(sms, dout); run_rate<OP_FFMA_IADD>(sms, dout);
  run_pair<V_GENERIC>(sms, g, dout2); run_pair<V_WRAPPED>(sms, g, dout2); run_pair<V_TWOIMG>(sms, g, dout2);
  run_pair<V_NOPBC>(sms, g, dout2); run_pair<V_GENERIC_X2>(sms, g, dout2); run_pair<V_TWOIMG_X2>(sms, g, dout2);
  run_pair<V_NOPBC_X2>(sms, g, dout2); run_pair<V_TRI>(sms, g, dout2);
  return 0;
}
