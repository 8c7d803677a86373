// PCIe probe for the host-resident step: which way of moving the rows the path reads is fastest?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/pcie_probe.cu -o gpurun_out/pcie_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// one warp per env: gather the needed rows of one env from host memory into packed device tensors
__global__ void gather_kernel(const float* __restrict__ dof, const float* __restrict__ root, const float* __restrict__ body,
                              const float* __restrict__ dforce, const float* __restrict__ ft,
                              float* __restrict__ d_dof, float* __restrict__ d_root, float* __restrict__ d_body,
                              float* __restrict__ d_dforce, float* __restrict__ d_ft, int N) {
  // flat element index over the 97 floats an env needs: 18 dof | 13 obj root | 39 tips | 9 dof force | 18 ft
  const long long total = (long long)N * 97;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int env = (int)(i / 97), c = (int)(i % 97);
    if (c < 18) d_dof[env * 18 + c] = dof[env * 18 + c];
    else if (c < 31) d_root[(env * 4 + 2) * 13 + c - 18] = root[(env * 4 + 2) * 13 + c - 18];
    else if (c < 70) { const int k = (c - 31) / 13, j = (c - 31) % 13; const int b = 6 + 5 * k; d_body[(env * 20 + b) * 13 + j] = body[(env * 20 + b) * 13 + j]; }
    else if (c < 79) d_dforce[env * 9 + c - 70] = dforce[env * 9 + c - 70];
    else d_ft[env * 18 + c - 79] = ft[env * 18 + c - 79];
  }
}


// rows only: the object root row and the three fingertip rows of each env (4 x 13 floats), read from host memory
__global__ void gather_rows(const float* __restrict__ root, const float* __restrict__ body,
                            float* __restrict__ d_root, float* __restrict__ d_body, int N) {
  const long long total = (long long)N * 52;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int env = (int)(i / 52), c = (int)(i % 52), k = c / 13, j = c % 13;
    if (k == 0) d_root[(env * 4 + 2) * 13 + j] = root[(env * 4 + 2) * 13 + j];
    else { const int b = 1 + 5 * k; d_body[(env * 20 + b) * 13 + j] = body[(env * 20 + b) * 13 + j]; }
  }
}
// same rows, fetched as whole 32-byte sectors: lane group of 4 x float4... here 8 lanes x float per sector
__global__ void gather_rows_sectors(const float* __restrict__ root, const float* __restrict__ body,
                                    float* __restrict__ d_root, float* __restrict__ d_body, int N) {
  // each row (52 B at a 4-byte-aligned address) is covered by 3 aligned 32-byte sectors = 24 floats; 24 lanes per row
  const long long total = (long long)N * 4 * 24;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 24; const int l = (int)(i % 24);
    const int env = (int)(r >> 2), k = (int)(r & 3);
    const float* src = k == 0 ? root + (env * 4 + 2) * 13 : body + (env * 20 + 1 + 5 * k) * 13;
    float* dst = k == 0 ? d_root + (env * 4 + 2) * 13 : d_body + (env * 20 + 1 + 5 * k) * 13;
    const long long base = ((long long)(size_t)src & ~31ll);
    const float* p = (const float*)base + l;
    const long long off = p - src;
    if (off >= 0 && off < 13) dst[off] = *p;
  }
}

// flat float4 copy from host-mapped memory (upper bound for zero-copy reads)
__global__ void zc_copy(const float4* __restrict__ src, float4* __restrict__ dst, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

template <class F> float timeit(cudaStream_t st, int reps, F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaStreamSynchronize(st));
  cudaEventRecord(a, st);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps * 1000.f;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 16384;
  float *h_dof, *h_root, *h_body, *h_df, *h_ft, *h_out;
  CK(cudaMallocHost(&h_dof, N * 18 * 4)); CK(cudaMallocHost(&h_root, N * 52 * 4)); CK(cudaMallocHost(&h_body, N * 260 * 4));
  CK(cudaMallocHost(&h_df, N * 9 * 4)); CK(cudaMallocHost(&h_ft, N * 18 * 4)); CK(cudaMallocHost(&h_out, (size_t)N * 160 * 4));
  memset(h_body, 1, N * 260 * 4);
  float *d_dof, *d_root, *d_body, *d_df, *d_ft, *d_out;
  CK(cudaMalloc(&d_dof, N * 18 * 4)); CK(cudaMalloc(&d_root, N * 52 * 4)); CK(cudaMalloc(&d_body, N * 260 * 4));
  CK(cudaMalloc(&d_df, N * 9 * 4)); CK(cudaMalloc(&d_ft, N * 18 * 4)); CK(cudaMalloc(&d_out, (size_t)N * 160 * 4));
  cudaStream_t st, st2; CK(cudaStreamCreate(&st)); CK(cudaStreamCreate(&st2));
  const cudaMemcpyKind HD = cudaMemcpyHostToDevice, DH = cudaMemcpyDeviceToHost;
  const int R = 20;
  auto small = [&] {
    cudaMemcpyAsync(d_dof, h_dof, N * 72, HD, st); cudaMemcpyAsync(d_df, h_df, N * 36, HD, st); cudaMemcpyAsync(d_ft, h_ft, N * 72, HD, st);
  };
  printf("N=%d\n", N);
  printf("small (dof+dforce+ft, %.2f MB): %.1f us\n", N * 180e-6, timeit(st, R, small));
  printf("root full (%.2f MB): %.1f us\n", N * 208e-6, timeit(st, R, [&] { cudaMemcpyAsync(d_root, h_root, N * 208, HD, st); }));
  printf("root 2D obj row only (52/208): %.1f us\n", timeit(st, R, [&] { cudaMemcpy2DAsync(d_root + 26, 208, h_root + 26, 208, 52, N, HD, st); }));
  printf("body full (%.2f MB): %.1f us\n", N * 1040e-6, timeit(st, R, [&] { cudaMemcpyAsync(d_body, h_body, N * 1040, HD, st); }));
  printf("body 2D 6..16 (572/1040, %.2f MB): %.1f us\n", N * 572e-6, timeit(st, R, [&] { cudaMemcpy2DAsync(d_body + 78, 1040, h_body + 78, 1040, 572, N, HD, st); }));
  printf("body 3x 2D tips (52/1040): %.1f us\n", timeit(st, R, [&] {
    for (int k = 0; k < 3; ++k) cudaMemcpy2DAsync(d_body + (6 + 5 * k) * 13, 1040, h_body + (6 + 5 * k) * 13, 1040, 52, N, HD, st); }));
  printf("body 1x 2D tips as 3N rows (52/260): %.1f us\n", timeit(st, R, [&] {
    cudaMemcpy2DAsync(d_body + 78, 260, h_body + 78, 260, 52, (size_t)N * 4 - 1, HD, st); }));
  for (int blocks : {148, 296, 592, 1184}) {
    printf("gather kernel (388 B/env) grid %d: %.1f us\n", blocks, timeit(st, R, [&] {
      gather_kernel<<<blocks, 512, 0, st>>>(h_dof, h_root, h_body, h_df, h_ft, d_dof, d_root, d_body, d_df, d_ft, N); }));
  }
  for (int blocks : {148, 592}) {
    const long long n4 = (long long)N * 260 / 4;
    printf("zero-copy flat read of body (%.2f MB) grid %d: %.1f us\n", N * 1040e-6, blocks, timeit(st, R, [&] {
      zc_copy<<<blocks, 512, 0, st>>>((const float4*)h_body, (float4*)d_body, n4); }));
  }
  const size_t outB = (size_t)N * 620;
  printf("D2H contiguous (%.2f MB): %.1f us\n", outB * 1e-6, timeit(st, R, [&] { cudaMemcpyAsync(h_out, d_out, outB, DH, st); }));
  for (int blocks : {148, 592}) {
    const long long n4 = (long long)outB / 16;
    printf("zero-copy flat write to host (%.2f MB) grid %d: %.1f us\n", outB * 1e-6, blocks, timeit(st, R, [&] {
      zc_copy<<<blocks, 512, 0, st>>>((const float4*)d_out, (float4*)h_out, n4); }));
  }
  // duplex: current upload on st, download on st2, concurrently
  {
    cudaEvent_t a, b, c; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
    auto up = [&] { small(); cudaMemcpyAsync(d_root, h_root, N * 208, HD, st); cudaMemcpy2DAsync(d_body + 78, 1040, h_body + 78, 1040, 572, N, HD, st); };
    printf("upload as shipped (960 B/env): %.1f us\n", timeit(st, R, up));
    cudaDeviceSynchronize();
    cudaEventRecord(a, st); cudaStreamWaitEvent(st2, a, 0);
    for (int i = 0; i < R; ++i) { up(); cudaMemcpyAsync(h_out, d_out, outB, DH, st2); }
    cudaEventRecord(c, st2); cudaStreamWaitEvent(st, c, 0); cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("duplex upload+download: %.1f us per pair\n", ms / R * 1000.f);
    cudaDeviceSynchronize();
    cudaEventRecord(a, st); cudaStreamWaitEvent(st2, a, 0);
    for (int i = 0; i < R; ++i) {
      gather_kernel<<<592, 512, 0, st>>>(h_dof, h_root, h_body, h_df, h_ft, d_dof, d_root, d_body, d_df, d_ft, N);
      cudaMemcpyAsync(h_out, d_out, outB, DH, st2);
    }
    cudaEventRecord(c, st2); cudaStreamWaitEvent(st, c, 0); cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&ms, a, b);
    printf("duplex gather+download: %.1f us per pair\n", ms / R * 1000.f);
  }
  // leaner upload: only the rows the path reads, spread over several copy engines
  {
    cudaStream_t s3, s4; CK(cudaStreamCreate(&s3)); CK(cudaStreamCreate(&s4));
    cudaEvent_t a, b, e2, e3, e4; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&e2); cudaEventCreate(&e3); cudaEventCreate(&e4);
    auto lean1 = [&] { small(); cudaMemcpy2DAsync(d_root + 26, 208, h_root + 26, 208, 52, N, HD, st);
      for (int k = 0; k < 3; ++k) cudaMemcpy2DAsync(d_body + (6 + 5 * k) * 13, 1040, h_body + (6 + 5 * k) * 13, 1040, 52, N, HD, st); };
    printf("lean upload, one stream: %.1f us\n", timeit(st, R, lean1));
    for (int ways = 2; ways <= 4; ++ways) {
      cudaDeviceSynchronize();
      cudaEventRecord(a, st);
      for (int i = 0; i < R; ++i) {
        cudaEventRecord(e2, st); cudaStreamWaitEvent(st2, e2, 0); cudaStreamWaitEvent(s3, e2, 0); cudaStreamWaitEvent(s4, e2, 0);
        cudaStream_t q[4] = {st, st2, s3, s4};
        small();
        cudaMemcpy2DAsync(d_root + 26, 208, h_root + 26, 208, 52, N, HD, q[1 % ways]);
        for (int k = 0; k < 3; ++k) cudaMemcpy2DAsync(d_body + (6 + 5 * k) * 13, 1040, h_body + (6 + 5 * k) * 13, 1040, 52, N, HD, q[(k + 1) % ways]);
        cudaEventRecord(e2, st2); cudaEventRecord(e3, s3); cudaEventRecord(e4, s4);
        cudaStreamWaitEvent(st, e2, 0); cudaStreamWaitEvent(st, e3, 0); cudaStreamWaitEvent(st, e4, 0);
      }
      cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("lean upload over %d streams: %.1f us\n", ways, ms / R * 1000.f);
    }
    // lean upload (2 streams) concurrent with the download
    cudaDeviceSynchronize();
    cudaEventRecord(a, st);
    for (int i = 0; i < R; ++i) {
      cudaEventRecord(e2, st); cudaStreamWaitEvent(st2, e2, 0); cudaStreamWaitEvent(s3, e2, 0);
      small(); cudaMemcpy2DAsync(d_root + 26, 208, h_root + 26, 208, 52, N, HD, st);
      for (int k = 0; k < 3; ++k) cudaMemcpy2DAsync(d_body + (6 + 5 * k) * 13, 1040, h_body + (6 + 5 * k) * 13, 1040, 52, N, HD, st2);
      cudaMemcpyAsync(h_out, d_out, outB, DH, s3);
      cudaEventRecord(e2, st2); cudaEventRecord(e3, s3); cudaStreamWaitEvent(st, e2, 0); cudaStreamWaitEvent(st, e3, 0);
    }
    cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("duplex lean(2 streams)+download: %.1f us per pair\n", ms / R * 1000.f);
    // download split in two halves on two streams
    cudaDeviceSynchronize();
    cudaEventRecord(a, st);
    for (int i = 0; i < R; ++i) {
      cudaEventRecord(e2, st); cudaStreamWaitEvent(st2, e2, 0);
      cudaMemcpyAsync(h_out, d_out, outB / 2, DH, st); cudaMemcpyAsync((char*)h_out + outB / 2, (char*)d_out + outB / 2, outB / 2, DH, st2);
      cudaEventRecord(e2, st2); cudaStreamWaitEvent(st, e2, 0);
    }
    cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&ms, a, b);
    printf("download split over 2 streams: %.1f us\n", ms / R * 1000.f);
  }
  {
    cudaEvent_t a, b, e2; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&e2);
    for (int blocks : {148, 592, 2368}) {
      printf("gather rows only (208 B/env) grid %d: %.1f us\n", blocks, timeit(st, R, [&] { gather_rows<<<blocks, 256, 0, st>>>(h_root, h_body, d_root, d_body, N); }));
      printf("gather rows by sectors grid %d: %.1f us\n", blocks, timeit(st, R, [&] { gather_rows_sectors<<<blocks, 256, 0, st>>>(h_root, h_body, d_root, d_body, N); }));
    }
    cudaDeviceSynchronize();
    cudaEventRecord(a, st);
    for (int i = 0; i < R; ++i) {
      cudaEventRecord(e2, st); cudaStreamWaitEvent(st2, e2, 0);
      small();
      gather_rows<<<592, 256, 0, st2>>>(h_root, h_body, d_root, d_body, N);
      cudaEventRecord(e2, st2); cudaStreamWaitEvent(st, e2, 0);
    }
    cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("small DMA || gather rows: %.1f us\n", ms / R * 1000.f);
  }
  return 0;
}
