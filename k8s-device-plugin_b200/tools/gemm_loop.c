/*
 * gemm_loop.c — BASELINE.json configs[3] / SURVEY.md §8d cfg 4: a cuBLAS SGEMM loop (default 8192^3) for a fixed
 * wall time, as an ordinary CUDA *runtime* application (cudart + cuBLAS resolve the driver through
 * cuGetProcAddress/dlsym, i.e. through the hook's symbol routing). It reports its own achieved GPU duty cycle:
 * sum of per-GEMM device time (cudaEvent pairs) / wall time — to be compared with CUDA_DEVICE_SM_LIMIT.
 */
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { fprintf(stderr, "gemm_loop: %s -> %s\n", #x, cudaGetErrorString(_e)); printf("{\"error\": \"%s\"}\n", cudaGetErrorString(_e)); return 3; } } while (0)
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 8192;
    double seconds = argc > 2 ? atof(argv[2]) : 10.0;
    float *a, *b, *c;
    size_t bytes = (size_t)n * n * sizeof(float);
    CK(cudaMalloc((void **)&a, bytes)); CK(cudaMalloc((void **)&b, bytes)); CK(cudaMalloc((void **)&c, bytes));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    cublasHandle_t h;
    if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) { printf("{\"error\": \"cublasCreate\"}\n"); return 3; }
    const float one = 1.f, zero = 0.f;
    enum { RING = 64 };
    cudaEvent_t e0[RING], e1[RING];
    for (int i = 0; i < RING; i++) { CK(cudaEventCreate(&e0[i])); CK(cudaEventCreate(&e1[i])); }
    for (int i = 0; i < 3; i++) cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, a, n, b, n, &zero, c, n);
    CK(cudaDeviceSynchronize());
    double t0 = now_s(), busy_ms = 0, max_ms = 0;
    long iters = 0;
    while (now_s() - t0 < seconds) {
        int k = (int)(iters % RING);
        if (iters >= RING) { float ms; CK(cudaEventSynchronize(e1[k])); CK(cudaEventElapsedTime(&ms, e0[k], e1[k])); busy_ms += ms; if (ms > max_ms) max_ms = ms; }
        CK(cudaEventRecord(e0[k], 0));
        if (cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, a, n, b, n, &zero, c, n) != CUBLAS_STATUS_SUCCESS) { printf("{\"error\": \"sgemm\"}\n"); return 3; }
        CK(cudaEventRecord(e1[k], 0));
        iters++;
    }
    CK(cudaDeviceSynchronize());
    double wall = now_s() - t0;
    long first = iters > RING ? iters - RING : 0;
    for (long i = first; i < iters; i++) { float ms; int k = (int)(i % RING); CK(cudaEventElapsedTime(&ms, e0[k], e1[k])); busy_ms += ms; if (ms > max_ms) max_ms = ms; }
    printf("{\"n\": %d, \"gemms\": %ld, \"wall_s\": %.3f, \"busy_s\": %.3f, \"duty\": %.4f, \"tflops\": %.1f, \"max_gemm_ms\": %.2f}\n",
           n, iters, wall, busy_ms / 1e3, busy_ms / 1e3 / wall, 2.0 * n * n * n * iters / wall / 1e12, max_ms);
    return 0;
}
