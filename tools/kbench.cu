// kbench.cu -- C-side micro-benchmark of the step kernel through the C ABI (no Python in the loop).
//   nvcc -O2 -o gpurun_out/kbench tools/kbench.cu -Ldcd_isaac_b200 -lmgplr -Xlinker -rpath=$PWD/dcd_isaac_b200
// usage: kbench N W T reps ablate   (ablate bits: 1 no image, 2 no scalar outputs, 4 no masks)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/mgplr.h"
extern "C" int mgplr_debug_prof(mgplr_venv *v, unsigned long long *out);

#define CKC(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
#define CKM(x) do { int rc = (x); if (rc) { printf("mgplr rc=%d %s line %d\n", rc, mgplr_last_error(), __LINE__); exit(1);} } while (0)

int main(int argc, char **argv) {
  int N = argc > 1 ? atoi(argv[1]) : 131072, W = argc > 2 ? atoi(argv[2]) : 15, T = argc > 3 ? atoi(argv[3]) : 64;
  int reps = argc > 4 ? atoi(argv[4]) : 5, ablate = argc > 5 ? atoi(argv[5]) : 0, see = argc > 6 ? atoi(argv[6]) : 1, amode = argc > 7 ? atoi(argv[7]) : 0, rr = argc > 8 ? atoi(argv[8]) : 0;
  mgplr_env_config cfg = {W, 5, 250, 250, see, 50, 0, 1, 0, 4};
  mgplr_venv *v;
  CKM(mgplr_venv_create(&cfg, N, 0, &v));
  std::vector<uint32_t> limbs(2 * N); std::vector<int32_t> cnt(N, 2);
  for (int i = 0; i < N; i++) { limbs[2 * i] = 12345u + 977u * i; limbs[2 * i + 1] = 99u + i; }
  CKM(mgplr_seed(v, limbs.data(), cnt.data(), nullptr, N, 0));
  {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CKM(mgplr_reset_random(v, nullptr, nullptr, 0)); CKC(cudaDeviceSynchronize());
    cudaEventRecord(a, 0);
    CKM(mgplr_reset_random(v, nullptr, nullptr, 0));
    cudaEventRecord(b, 0); CKC(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("reset_random of %d envs: %.1f us (%.1f ns/env)\n", N, ms * 1e3, ms * 1e6 / N);
  }
  float *img, *dir, *rew, *mk, *bm, *cm; uint8_t *fl; int64_t *act; float *epr; int32_t *epl;
  CKC(cudaMalloc(&img, (size_t)(T + 1) * N * 300)); CKC(cudaMalloc(&dir, (size_t)(T + 1) * N * 4)); CKC(cudaMalloc(&rew, (size_t)T * N * 4));
  CKC(cudaMalloc(&mk, (size_t)(T + 1) * N * 4)); CKC(cudaMalloc(&bm, (size_t)(T + 1) * N * 4)); CKC(cudaMalloc(&cm, (size_t)(T + 1) * N * 4));
  CKC(cudaMalloc(&fl, (size_t)T * N)); CKC(cudaMalloc(&act, (size_t)T * N * 8)); CKC(cudaMalloc(&epr, N * 4)); CKC(cudaMalloc(&epl, N * 4));
  std::vector<int64_t> h((size_t)T * N);
  srand(1);
  for (auto &x : h) x = amode == 1 ? (rand() & 1) : ((rand() & 1) ? 2 : rand() % 7);
  CKC(cudaMemcpy(act, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  cudaStream_t st; CKC(cudaStreamCreate(&st));
  auto steps = [&]() {
    mgplr_step_out ro = {}; ro.image = img; ro.direction = dir;
    CKM(mgplr_reset_agent(v, &ro, st));  // start of a rollout (also replays deferred respawn draws)
    for (int t = 0; t < T; t++) {
      mgplr_step_out o = {};
      if (!(ablate & 1)) o.image = img + (size_t)(t + 1) * N * 75;
      if (!(ablate & 2)) { o.direction = dir + (size_t)(t + 1) * N; o.reward = rew + (size_t)t * N; o.flags = fl + (size_t)t * N; o.ep_return = epr; o.ep_length = epl; }
      if (!(ablate & 4)) { o.masks = mk + (size_t)(t + 1) * N; o.bad_masks = bm + (size_t)(t + 1) * N; o.cliffhanger_masks = cm + (size_t)(t + 1) * N; }
      CKM(mgplr_step_env(v, act + (size_t)t * N, rr, nullptr, 0, &o, st));
    }
  };
  steps(); CKC(cudaStreamSynchronize(st));
  if (getenv("KB_TRACE")) {  // per-launch timing of one rollout (no graph): print the slow launches
    std::vector<cudaEvent_t> ev(T + 2);
    for (auto &e : ev) cudaEventCreate(&e);
    mgplr_step_out ro = {}; ro.image = img; ro.direction = dir;
    cudaEventRecord(ev[0], st);
    CKM(mgplr_reset_agent(v, &ro, st));
    cudaEventRecord(ev[1], st);
    for (int t = 0; t < T; t++) {
      mgplr_step_out o = {};
      o.image = img + (size_t)(t + 1) * N * 75; o.reward = rew + (size_t)t * N; o.flags = fl + (size_t)t * N;
      CKM(mgplr_step_env(v, act + (size_t)t * N, rr, nullptr, 0, &o, st));
      cudaEventRecord(ev[t + 2], st);
    }
    CKC(cudaStreamSynchronize(st));
    float ms0; cudaEventElapsedTime(&ms0, ev[0], ev[1]);
    printf("trace: reset_agent %.1f us\n", ms0 * 1e3);
    for (int t = 0; t < T; t++) {
      float msx; cudaEventElapsedTime(&msx, ev[t + 1], ev[t + 2]);
      if (msx * 1e3 > 80.0 || t < 2) printf("trace: step %d  %.1f us\n", t, msx * 1e3);
    }
  }
  if (getenv("MGPLR_RR_PROF")) {  // phase timing of single DR launches
    unsigned long long pr[24];
    mgplr_debug_prof(v, pr);
    for (int t = 0; t < 6; t++) {
      mgplr_step_out o = {};
      o.image = img + (size_t)(t + 1) * N * 75; o.reward = rew + (size_t)t * N; o.flags = fl + (size_t)t * N;
      CKM(mgplr_step_env(v, act + (size_t)t * N, rr, nullptr, 0, &o, st));
      CKC(cudaStreamSynchronize(st));
      mgplr_debug_prof(v, pr);
      printf("prof step %d: tiles first->last %.1f us, regen phase %.1f us, jobs %llu, avg %.1f us (tables %.1f, level %.1f), max %.1f us\n", t,
             (pr[1] - pr[0]) * 1e-3, pr[2] > pr[1] ? (pr[2] - pr[1]) * 1e-3 : 0.0, pr[4], pr[4] ? pr[3] / (double)pr[4] / 1965.0 : 0.0,
             pr[4] ? pr[6] / (double)pr[4] / 1965.0 : 0.0, pr[4] ? pr[7] / (double)pr[4] / 1965.0 : 0.0, pr[5] / 1965.0);
      if (pr[18]) printf("   tiles with a commit (%llu): step block %.2f us, commit loop %.2f us, rebuild+render+emit %.2f us, write-back %.2f us\n", pr[18],
             pr[16] / (double)pr[18] / 1965.0, pr[17] / (double)pr[18] / 1965.0, pr[19] / (double)pr[18] / 1965.0, pr[20] / (double)pr[18] / 1965.0);
      printf("   resets from candidates %llu, rebuilt in the kernel %llu; tile phase per warp: no reset avg %.1f max %.1f us (%llu warps), with resets avg %.1f max %.1f us (%llu warps)\n", pr[8], pr[9],
             pr[11] ? pr[10] / (double)pr[11] * 1e-3 : 0.0, pr[12] * 1e-3, pr[11], pr[14] ? pr[13] / (double)pr[14] * 1e-3 : 0.0, pr[15] * 1e-3, pr[14]);
    }
  }
  cudaGraph_t g; cudaGraphExec_t ge;
  CKC(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal)); steps(); CKC(cudaStreamEndCapture(st, &g));
  CKC(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  CKC(cudaGraphLaunch(ge, st)); CKC(cudaStreamSynchronize(st));
  cudaEventRecord(a, st);
  for (int r = 0; r < reps; r++) CKC(cudaGraphLaunch(ge, st));
  cudaEventRecord(b, st); CKC(cudaStreamSynchronize(st));
  float ms; cudaEventElapsedTime(&ms, a, b);
  double us = ms * 1e3 / (reps * (double)T);  // includes 1/T of the reset_agent launch
  printf("amode=%d rr=%d ", amode, rr); printf("N=%d W=%d T=%d see=%d ablate=%d tile=%s: %.2f us/launch  %.3f Gsteps/s  %.0f GB/s@360B\n", N, W, T, see, ablate,
         getenv("MGPLR_TILE") ? getenv("MGPLR_TILE") : "def", us, N / us * 1e-3, N * 360.0 / us * 1e-3);
  {  // the multi-step kernel: T transitions in one launch from a recorded u8 action stream
    std::vector<uint8_t> h8((size_t)T * N);
    for (size_t i = 0; i < h8.size(); i++) h8[i] = (uint8_t)h[i];
    uint8_t *act8; CKC(cudaMalloc(&act8, h8.size())); CKC(cudaMemcpy(act8, h8.data(), h8.size(), cudaMemcpyHostToDevice));
    mgplr_step_out o = {};
    o.image = img + (size_t)N * 75; o.direction = dir + N; o.reward = rew; o.flags = fl; o.masks = mk + N; o.bad_masks = bm + N;
    o.cliffhanger_masks = cm + N;
    CKM(mgplr_rollout(v, act8, T, rr, &o, st)); CKC(cudaStreamSynchronize(st));
    cudaEventRecord(a, st);
    for (int r = 0; r < reps; r++) CKM(mgplr_rollout(v, act8, T, rr, &o, st));
    cudaEventRecord(b, st); CKC(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ms, a, b);
    double us2 = ms * 1e3 / (reps * (double)T);
    printf("  mgplr_rollout (T steps per launch): %.2f us/step  %.3f Gsteps/s  %.0f GB/s@360B\n", us2, N / us2 * 1e-3, N * 360.0 / us2 * 1e-3);
  }
  mgplr_venv_destroy(v);
  return 0;
}
