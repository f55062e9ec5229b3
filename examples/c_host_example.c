/* The QLoRA linear hot path driven from plain C through the C ABI only -- no Python, no torch.
 *
 *   gcc -std=c11 -O2 -Iinclude -I/usr/local/cuda/include examples/c_host_example.c \
 *       -Lcausal-unified-language-vision_b200 -lb2q -L/usr/local/cuda/lib64 -lcudart -lm -o c_host_example
 *   LD_LIBRARY_PATH=causal-unified-language-vision_b200 ./c_host_example
 *
 * What a non-Python host of the reference's path (a C++ trainer, a serving runtime) would do: quantise a weight to the
 * bitsandbytes NF4 layout, run forward and backward of one QLoRA linear, and check a sample of the outputs against a
 * double-precision host sum over the weights the library itself decodes.  Exit code 0 = all checks within 2e-2.
 * tests/test_gpu_c_host.py builds and runs it on a B200; the CPU suite only compiles and links it. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2q.h"

static const float NF4[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                              -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                              0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                              0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

static uint16_t f2bf(float f) { /* round to nearest even */
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static uint32_t rng_state = 12345u;
static float urand(void) { /* uniform in (-1, 1) */
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) / 8388608.0f) - 1.0f;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        int rc_ = (call);                                                                               \
        if (rc_ != 0) {                                                                                 \
            fprintf(stderr, "%s -> %d (%s) %s\n", #call, rc_, b2q_error_string(rc_), b2q_last_error_detail()); \
            return 2;                                                                                   \
        }                                                                                               \
    } while (0)
#define CU(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s -> %s\n", #call, cudaGetErrorString(e_));              \
            return 2;                                                                  \
        }                                                                              \
    } while (0)

static void* dev_bf16(const float* src, size_t n) { /* host fp32 -> device bf16 */
    uint16_t* h = (uint16_t*)malloc(n * 2);
    void* d = NULL;
    for (size_t i = 0; i < n; ++i) h[i] = f2bf(src[i]);
    if (cudaMalloc(&d, n * 2) != cudaSuccess) return NULL;
    cudaMemcpy(d, h, n * 2, cudaMemcpyHostToDevice);
    free(h);
    return d;
}
static float* host_from_bf16(const void* d, size_t n) { /* device bf16 -> host fp32 */
    uint16_t* h = (uint16_t*)malloc(n * 2);
    float* f = (float*)malloc(n * 4);
    cudaMemcpy(h, d, n * 2, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n; ++i) f[i] = bf2f(h[i]);
    free(h);
    return f;
}

int main(void) {
    const int M = 384, N = 512, K = 768, r = 64;
    const float s = 16.0f / r;
    const size_t nW = (size_t)N * K;
    float *W = malloc(nW * 4), *x = malloc((size_t)M * K * 4), *dy = malloc((size_t)M * N * 4);
    float *A = malloc((size_t)r * K * 4), *B = malloc((size_t)N * r * 4);
    for (size_t i = 0; i < nW; ++i) W[i] = 0.02f * urand();
    for (size_t i = 0; i < (size_t)M * K; ++i) x[i] = bf2f(f2bf(urand()));
    for (size_t i = 0; i < (size_t)M * N; ++i) dy[i] = bf2f(f2bf(urand() / sqrtf((float)N)));
    for (size_t i = 0; i < (size_t)r * K; ++i) A[i] = bf2f(f2bf(urand() / sqrtf((float)K)));
    for (size_t i = 0; i < (size_t)N * r; ++i) B[i] = bf2f(f2bf(0.02f * urand()));

    printf("b2q version %d\n", b2q_version());
    /* ---- quantise W to NF4 (blocksize 64, plain fp32 absmax) on the device ---- */
    float *dW, *d_absmax, *d_code;
    uint8_t* d_packed;
    CU(cudaMalloc((void**)&dW, nW * 4));
    CU(cudaMemcpy(dW, W, nW * 4, cudaMemcpyHostToDevice));
    CU(cudaMalloc((void**)&d_packed, nW / 2));
    CU(cudaMalloc((void**)&d_absmax, nW / 64 * 4));
    CU(cudaMalloc((void**)&d_code, 64));
    CU(cudaMemcpy(d_code, NF4, 64, cudaMemcpyHostToDevice));
    CK(b2q_nf4_quantize(dW, 0, (int64_t)nW, d_packed, d_absmax, 0));
    b2q_nf4_weight w;
    memset(&w, 0, sizeof w);
    w.packed = d_packed;
    w.absmax = d_absmax;
    w.code16 = d_code;
    /* the weights the kernels will see, decoded by the library (bit-exact with the in-kernel decode) */
    void* d_wdec;
    CU(cudaMalloc(&d_wdec, nW * 2));
    CK(b2q_nf4_decode(d_packed, d_absmax, NULL, NULL, NULL, 0.f, d_code, d_wdec, (int64_t)nW, 64, 1, 0));

    /* ---- forward + backward of one QLoRA linear ---- */
    void *dx_in = dev_bf16(x, (size_t)M * K), *d_dy = dev_bf16(dy, (size_t)M * N);
    void *dA_w = dev_bf16(A, (size_t)r * K), *dB_w = dev_bf16(B, (size_t)N * r);
    void *d_u, *d_us, *d_y, *d_du, *d_dx, *d_gA, *d_gB, *d_ws;
    CU(cudaMalloc(&d_u, (size_t)M * r * 2));
    CU(cudaMalloc(&d_us, (size_t)M * r * 2));
    CU(cudaMalloc(&d_y, (size_t)M * N * 2));
    CU(cudaMalloc(&d_du, (size_t)M * r * 2));
    CU(cudaMalloc(&d_dx, (size_t)M * K * 2));
    CU(cudaMalloc(&d_gA, (size_t)r * K * 2));
    CU(cudaMalloc(&d_gB, (size_t)N * r * 2));
    const size_t ws_bytes = b2q_lora_grads_workspace_bytes(M, N, K, r);
    CU(cudaMalloc(&d_ws, ws_bytes));
    const uint64_t launches0 = b2q_launch_count();
    CK(b2q_lora_down(dx_in, dA_w, s, 0, 0.f, d_u, d_us, M, K, r, 0));
    CK(b2q_qlora_fwd(dx_in, &w, d_us, dB_w, d_y, M, N, K, r, 0));
    CK(b2q_lora_bwd_du(d_dy, dB_w, s, 0.f, d_du, M, N, r, 0));
    CK(b2q_qlora_bwd_dx(d_dy, &w, d_du, dA_w, 0, 0.f, d_dx, M, N, K, r, 0));
    CK(b2q_lora_grads(d_dy, dx_in, d_u, d_du, s, 0, 0.f, d_gA, d_gB, 0, d_ws, ws_bytes, M, N, K, r, 0));
    CU(cudaDeviceSynchronize());
    printf("kernel launches: %llu\n", (unsigned long long)(b2q_launch_count() - launches0));

    /* ---- check a sample of y, dx, dA, dB against double-precision host sums ---- */
    float *Wd = host_from_bf16(d_wdec, nW), *y = host_from_bf16(d_y, (size_t)M * N), *dxo = host_from_bf16(d_dx, (size_t)M * K);
    float *gA = host_from_bf16(d_gA, (size_t)r * K), *gB = host_from_bf16(d_gB, (size_t)N * r);
    float *u = host_from_bf16(d_u, (size_t)M * r), *du = host_from_bf16(d_du, (size_t)M * r);
    double worst = 0.0, scale_y = 0.0, scale_dx = 0.0, scale_gA = 0.0, scale_gB = 0.0;
    for (size_t i = 0; i < (size_t)M * N; ++i) scale_y = fmax(scale_y, fabs(y[i]));
    for (size_t i = 0; i < (size_t)M * K; ++i) scale_dx = fmax(scale_dx, fabs(dxo[i]));
    for (size_t i = 0; i < (size_t)r * K; ++i) scale_gA = fmax(scale_gA, fabs(gA[i]));
    for (size_t i = 0; i < (size_t)N * r; ++i) scale_gB = fmax(scale_gB, fabs(gB[i]));
    for (int t = 0; t < 64; ++t) {
        const int m = (t * 37) % M, n = (t * 101) % N, k = (t * 53) % K, j = (t * 7) % r;
        double ref = 0.0, lora = 0.0;
        for (int kk = 0; kk < K; ++kk) ref += (double)x[(size_t)m * K + kk] * Wd[(size_t)n * K + kk];
        for (int jj = 0; jj < r; ++jj) lora += (double)bf2f(f2bf(s * u[(size_t)m * r + jj])) * B[(size_t)n * r + jj];
        worst = fmax(worst, fabs(ref + lora - y[(size_t)m * N + n]) / scale_y);
        ref = 0.0;
        for (int nn = 0; nn < N; ++nn) ref += (double)dy[(size_t)m * N + nn] * Wd[(size_t)nn * K + k];
        for (int jj = 0; jj < r; ++jj) ref += (double)du[(size_t)m * r + jj] * A[(size_t)jj * K + k];
        worst = fmax(worst, fabs(ref - dxo[(size_t)m * K + k]) / scale_dx);
        ref = 0.0;
        for (int mm = 0; mm < M; ++mm) ref += (double)du[(size_t)mm * r + j] * x[(size_t)mm * K + k];
        worst = fmax(worst, fabs(ref - gA[(size_t)j * K + k]) / scale_gA);
        ref = 0.0;
        for (int mm = 0; mm < M; ++mm) ref += (double)dy[(size_t)mm * N + n] * u[(size_t)mm * r + j];
        worst = fmax(worst, fabs(s * ref - gB[(size_t)n * r + j]) / scale_gB);
    }
    printf("worst sampled error relative to the tensor's max: %.3e\n", worst);
    if (!(worst <= 2e-2)) {
        fprintf(stderr, "FAILED\n");
        return 1;
    }
    printf("C host example OK\n");
    return 0;
}
