// launch_latency.cu — microbenchmark: what the SUBMISSION of the batch-1 search costs, with the kernels' own work taken out.
// Two stand-in kernels with the shapes of the latency path (kA = K1: 148 CTAs x 512 threads, reads the 6 KB query;
// kB = the fused K3+K4+K5 tail: 13 CTAs x 160 threads, last CTA writes ~300 B of results into pinned host memory), each
// spinning `work` cycles, submitted the ways the library could submit them:
//   A  graph [H2D 6.2 KB -> kA -> kB] + cudaStreamSynchronize              (round 2's path)
//   B  the same graph, completion seen by polling a sequence word in pinned memory
//   C  graph with a PROGRAMMATIC edge kA -> kB (kB resident early, griddepcontrol.wait) + polling
//   D  no graph: cudaMemcpyAsync + kA + kB (programmatic stream serialization) + polling
//   E  no graph, no copy: the query travels as a 6 KB __grid_constant__ kernel parameter, kB programmatic + polling
//   F  graph [kA' -> kB] where CTA 0 of kA' pulls the query from pinned host memory and publishes it (no copy node) + polling
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_latency launch_latency.cu && ./launch_latency
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int QN = 1536;
struct qparam { float q[QN]; unsigned seq; };

__device__ __forceinline__ void spin(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

// MODE 0: query from device memory; 1: from the kernel parameter; 2: CTA 0 pulls it from host memory and publishes it
template <int MODE>
__global__ void __launch_bounds__(512) kA(const float* __restrict__ q_dev, const __grid_constant__ qparam P, const float* q_host,
                                          float* q_pub, unsigned* pub_flag, unsigned seq, float* partial, long long work, int pdl) {
  __shared__ float qs[QN];
  if (pdl) asm volatile("griddepcontrol.launch_dependents;");
  if (MODE == 0) for (int i = threadIdx.x; i < QN; i += 512) qs[i] = q_dev[i];
  if (MODE == 1) for (int i = threadIdx.x; i < QN; i += 512) qs[i] = P.q[i];
  if (MODE == 2) {
    if (blockIdx.x == 0) {
      for (int i = threadIdx.x; i < QN; i += 512) { const float v = q_host[i]; qs[i] = v; q_pub[i] = v; }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicExch(pub_flag, seq);
    } else {
      if (threadIdx.x == 0) while (atomicAdd(pub_flag, 0u) != seq) {}
      __syncthreads();
      for (int i = threadIdx.x; i < QN; i += 512) qs[i] = __ldcg(q_pub + i);
    }
  }
  __syncthreads();
  spin(work);
  float s = 0.f;
  for (int i = threadIdx.x; i < QN; i += 512) s += qs[i];
  if (threadIdx.x < 32) partial[blockIdx.x * 32 + threadIdx.x] = s;
}

__global__ void __launch_bounds__(160) kB(const float* __restrict__ partial, const unsigned* seq_dev, unsigned seq_arg, unsigned* ticket,
                                          float* out_host, volatile unsigned* flag_host, long long work, int pdl) {
  __shared__ int last;
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  float s = 0.f;
  for (int i = threadIdx.x; i < 148 * 32; i += 160) s += __ldcg(partial + i);
  spin(work);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    last = t == gridDim.x - 1;
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < 72) out_host[threadIdx.x] = s + threadIdx.x;   // ~290 bytes of results
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned seq = seq_dev ? __ldcg(seq_dev) : seq_arg;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_host), "r"(seq) : "memory");
  }
}

// G/H: ONE kernel — the 148 CTAs do kA's work, count themselves done, and the first 13 carry on with kB's work once all have
template <int MODE>
__global__ void __launch_bounds__(512) kAB(const float* __restrict__ q_dev, const __grid_constant__ qparam P, const unsigned* seq_dev,
                                           unsigned seq_arg, float* partial, unsigned* done, unsigned* ticket, float* out_host,
                                           volatile unsigned* flag_host, long long workA, long long workB) {
  __shared__ float qs[QN];
  __shared__ int last;
  if (MODE == 0) for (int i = threadIdx.x; i < QN; i += 512) qs[i] = q_dev[i];
  if (MODE == 1) for (int i = threadIdx.x; i < QN; i += 512) qs[i] = P.q[i];
  __syncthreads();
  spin(workA);
  float s = 0.f;
  for (int i = threadIdx.x; i < QN; i += 512) s += qs[i];
  if (threadIdx.x < 32) partial[blockIdx.x * 32 + threadIdx.x] = s;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(done, 1u);
  if (blockIdx.x >= 13 || threadIdx.x >= 160) return;
  if (threadIdx.x == 0) while (atomicAdd(done, 0u) < gridDim.x) {}
  asm volatile("bar.sync 1, 160;" ::: "memory");
  s = 0.f;
  for (int i = threadIdx.x; i < 148 * 32; i += 160) s += __ldcg(partial + i);
  spin(workB);
  __threadfence();
  asm volatile("bar.sync 1, 160;" ::: "memory");
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    last = t == 12;
    if (last) { *ticket = 0u; *done = 0u; }
  }
  asm volatile("bar.sync 1, 160;" ::: "memory");
  if (!last) return;
  if (threadIdx.x < 72) out_host[threadIdx.x] = s + threadIdx.x;
  __threadfence_system();
  asm volatile("bar.sync 1, 160;" ::: "memory");
  if (threadIdx.x == 0) {
    const unsigned seq = seq_dev ? __ldcg(seq_dev) : seq_arg;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_host), "r"(seq) : "memory");
  }
}
__global__ void kNull(volatile unsigned* flag_host, unsigned seq) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_host), "r"(seq) : "memory");
}

static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 3000;
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  float *h_in, *h_out, *d_in, *d_partial, *d_pub;
  unsigned *h_flag, *d_ticket, *d_pubflag, *d_done;
  const size_t in_bytes = QN * 4 + 96;  // query + keyword list + counts + the sequence word
  CK(cudaHostAlloc((void**)&h_in, 8192, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&h_out, 4096, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&h_flag, 64, cudaHostAllocDefault));
  CK(cudaMalloc((void**)&d_in, 8192));
  CK(cudaMalloc((void**)&d_partial, 148 * 32 * 4));
  CK(cudaMalloc((void**)&d_pub, 8192));
  CK(cudaMalloc((void**)&d_ticket, 64));
  CK(cudaMalloc((void**)&d_pubflag, 64));
  CK(cudaMalloc((void**)&d_done, 64));
  CK(cudaMemset(d_done, 0, 64));
  CK(cudaMemset(d_ticket, 0, 64));
  CK(cudaMemset(d_pubflag, 0, 64));
  memset(h_in, 0, 8192);
  *h_flag = 0;
  unsigned* seq_in_block = (unsigned*)((char*)h_in + QN * 4 + 92);
  const unsigned* d_seq = (const unsigned*)((char*)d_in + QN * 4 + 92);
  static qparam P;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);

  for (long long work_us : {0LL, 6LL}) {
    const long long workA = work_us * khz / 1000, workB = work_us * 3 * khz / 1000;  // kA ~6 us, kB ~18 us like C1
    auto launch_pair = [&](int modeA, bool pdl, unsigned seq, const unsigned* seq_dev) -> cudaError_t {
      if (modeA == 0) kA<0><<<148, 512, 0, st>>>(d_in, P, nullptr, nullptr, nullptr, seq, d_partial, workA, pdl);
      else if (modeA == 1) kA<1><<<148, 512, 0, st>>>(nullptr, P, nullptr, nullptr, nullptr, seq, d_partial, workA, pdl);
      else kA<2><<<148, 512, 0, st>>>(nullptr, P, h_in, d_pub, d_pubflag, seq, d_partial, workA, pdl);
      cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
      cfg.gridDim = dim3(13); cfg.blockDim = dim3(160); cfg.stream = st;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at.val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = &at; cfg.numAttrs = pdl ? 1 : 0;
      return cudaLaunchKernelEx(&cfg, kB, (const float*)d_partial, seq_dev, seq, d_ticket, h_out, (volatile unsigned*)h_flag, workB, pdl ? 1 : 0);
    };
    auto make_graph = [&](bool copy, int modeA, bool pdl, cudaGraphExec_t* exec) -> int {
      cudaGraph_t g;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      if (copy) CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, st));
      CK(launch_pair(modeA, pdl, 0, copy ? d_seq : nullptr));
      CK(cudaStreamEndCapture(st, &g));
      CK(cudaGraphInstantiate(exec, g, 0));
      cudaGraphDestroy(g);
      return 0;
    };
    cudaGraphExec_t gA, gC, gF;
    if (make_graph(true, 0, false, &gA)) return 1;
    if (make_graph(true, 0, true, &gC)) return 1;
    // F: the sequence word travels in the host block as well: kA' publishes it with the query (slot QN-1 of the block is unused)
    unsigned seq = 0;
    auto poll = [&](unsigned want) { while (*(volatile unsigned*)h_flag != want) {} };
    struct variant { const char* name; int id; };
    const variant vs[] = {{"A graph[H2D,kA,kB] + streamSync", 0}, {"B graph[H2D,kA,kB] + poll", 1}, {"C graph[H2D,kA,kB programmatic] + poll", 2},
                          {"D direct H2D,kA,kB(programmatic) + poll", 3}, {"E direct kA(param 6KB),kB(programmatic) + poll", 4},
                          {"E' direct kA(param 6KB),kB + poll", 5}, {"F direct kA(CTA0 pulls from host),kB(programmatic) + poll", 6},
                          {"G ONE kernel (param 6KB), hand-over through a counter + poll", 7}, {"H direct H2D + ONE kernel + poll", 8},
                          {"I null kernel (1 thread writes the flag) + poll", 9}, {"J null kernel + streamSync", 10},
                          {"K ONE kernel (param 6KB) + streamSync", 11}};
    printf("--- stand-in work: kA %lld us, kB %lld us ---\n", work_us, work_us * 3);
    for (const variant& v : vs) {
      std::vector<double> t(iters);
      for (int i = -50; i < iters; i++) {
        const double t0 = now_us();
        seq++;
        for (int j = 0; j < QN; j += 64) h_in[j] = (float)(seq + j);  // the caller's query lands in the pinned block
        *seq_in_block = seq;
        switch (v.id) {
          case 0: CK(cudaGraphLaunch(gA, st)); CK(cudaStreamSynchronize(st)); break;
          case 1: CK(cudaGraphLaunch(gA, st)); poll(seq); break;
          case 2: CK(cudaGraphLaunch(gC, st)); poll(seq); break;
          case 3: CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, st)); CK(launch_pair(0, true, seq, d_seq)); poll(seq); break;
          case 4: memcpy(P.q, h_in, QN * 4); CK(launch_pair(1, true, seq, nullptr)); poll(seq); break;
          case 5: memcpy(P.q, h_in, QN * 4); CK(launch_pair(1, false, seq, nullptr)); poll(seq); break;
          case 6: CK(launch_pair(2, true, seq, nullptr)); poll(seq); break;
          case 7: memcpy(P.q, h_in, QN * 4); kAB<1><<<148, 512, 0, st>>>(nullptr, P, nullptr, seq, d_partial, d_done, d_ticket, h_out, h_flag, workA, workB); poll(seq); break;
          case 8: CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, st)); kAB<0><<<148, 512, 0, st>>>(d_in, P, d_seq, seq, d_partial, d_done, d_ticket, h_out, h_flag, workA, workB); poll(seq); break;
          case 9: kNull<<<1, 1, 0, st>>>(h_flag, seq); poll(seq); break;
          case 10: kNull<<<1, 1, 0, st>>>(h_flag, seq); CK(cudaStreamSynchronize(st)); break;
          case 11: memcpy(P.q, h_in, QN * 4); kAB<1><<<148, 512, 0, st>>>(nullptr, P, nullptr, seq, d_partial, d_done, d_ticket, h_out, h_flag, workA, workB); CK(cudaStreamSynchronize(st)); break;
        }
        if (i >= 0) t[i] = now_us() - t0;
      }
      CK(cudaStreamSynchronize(st));
      std::sort(t.begin(), t.end());
      printf("%-62s p50 %6.2f us  p10 %6.2f  p90 %6.2f  p99 %6.2f\n", v.name, t[iters / 2], t[iters / 10], t[iters * 9 / 10], t[iters * 99 / 100]);
    }
    cudaGraphExecDestroy(gA); cudaGraphExecDestroy(gC);
    (void)gF;
  }
  return 0;
}
