// microbench.cu -- measured denominators next to the driver's HBM / bf16 peaks (MEASURED_PEAKS.json has neither FP64 nor shared memory):
//   fp64_fma   : DFMA issue peak (8 independent chains per thread)
//   dmma_884   : mma.sync.m8n8k4.f64      (the legacy FP64 tensor-pipe shape)
//   dmma_16816 : mma.sync.m16n8k16.f64    (sm_90+ shape)
//   smem_ld    : conflict-free 16-byte shared-memory loads
//   int_imad   : IMAD issue peak (LK's inner loop is integer-issue bound)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu ; prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s\"}\n", cudaGetErrorString(e_), #x); return 1; } } while (0)

__global__ void k_fma(double *out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; i++) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__global__ void k_dmma884(double *out, int iters, double a, double b) {
    double c[4][2];
    for (int q = 0; q < 4; q++) { c[q][0] = threadIdx.x + q; c[q][1] = q; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int q = 0; q < 4; q++) s += c[q][0] + c[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double *out, int iters, double a, double b) {
    double c[4][4];
    for (int q = 0; q < 4; q++) for (int r = 0; r < 4; r++) c[q][r] = threadIdx.x + q + r;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0; for (int q = 0; q < 4; q++) for (int r = 0; r < 4; r++) s += c[q][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_smem(double *out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x * 2;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double2 v = *reinterpret_cast<const double2 *>(sm + ((idx + u * 1024) & 8191));
            acc.x += v.x; acc.y += v.y;
        }
        idx = (idx + 2 * blockDim.x) & 8191;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

__global__ void k_imad(int *out, int iters, int a, int b) {
    int c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; i++) {
        c0 = c0 * a + b; c1 = c1 * a + b; c2 = c2 * a + b; c3 = c3 * a + b; c4 = c4 * a + b; c5 = c5 * a + b; c6 = c6 * a + b; c7 = c7 * a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

template <typename F> static float best_ms(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 8; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1); if (r >= 2 && t < best) best = t;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, nb = sms * 8, nt = 256, iters = 20000;
    double *out; CK(cudaMalloc(&out, (size_t)nb * nt * 8 * 2));
    const double thr = (double)nb * nt;
    float t_fma = best_ms([&] { k_fma<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    float t_884 = best_ms([&] { k_dmma884<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    float t_168 = best_ms([&] { k_dmma16816<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    float t_sm = best_ms([&] { k_smem<<<sms * 2, 512, 65536>>>(out, iters / 4); });
    float t_im = best_ms([&] { k_imad<<<nb, nt>>>((int *)out, iters, 3, 7); });
    CK(cudaDeviceSynchronize());
    const double fma_tf = thr * iters * 8 * 2 / (t_fma * 1e-3) / 1e12;
    const double d884_tf = (thr / 32) * iters * 4 * (2.0 * 8 * 8 * 4) / (t_884 * 1e-3) / 1e12;
    const double d168_tf = (thr / 32) * iters * 4 * (2.0 * 16 * 8 * 16) / (t_168 * 1e-3) / 1e12;
    const double sm_tb = (double)sms * 2 * 512 * (iters / 4) * 8 * 16 / (t_sm * 1e-3) / 1e12;
    const double im_tops = thr * iters * 8 / (t_im * 1e-3) / 1e12;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m16n8k16_tflops\": %.2f, \"smem_ld_tbs\": %.2f, \"imad_tops\": %.2f, "
           "\"how\": \"CUDA events, best of 6 after 2 warm-ups; %d blocks x %d threads, %d iterations x 8 independent chains\"}\n",
           p.name, sms, fma_tf, d884_tf, d168_tf, sm_tb, im_tops, nb, nt, iters);
    return 0;
}
