// zz_gpu.cu — sm_100a kernels and the C ABI (include/zigzag_gpu.h) of zigzag's per-gene MCMC sweep.
//
// Layout in HBM (per handle = one gene shard on one GPU):
//   Xg      double [R][G]        library-major (== R's column-major G x R matrix), shared by all chains
//   gl      double [G]           gene_lengths (unit mean)
//   per chain c (arrays are [C][G]): Yg, variance_g, XgLikelihood, variance_g_probability, tuningParam_yg,
//   tuningParam_variance_g (double), flags (uint16: active | in-spike | all-zero | pinned | k-1<<8),
//   win_y / win_v (2 x u64 accept windows, burn-in only), alloc_trace uint32 [C][K+1][G]
//   hyper   zz_hyper [C]         read by every block into shared memory
// One thread owns one gene; consecutive threads own consecutive genes so every array access is a coalesced
// 8-byte-per-lane stream; the gene's state stays in registers across the moves of a sweep.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <new>
#include <algorithm>
#include <thread>
#include "zz_device.cuh"
#include "zz_fast.cuh"
#include "zz_chain.cuh"

using namespace zz;

namespace {

constexpr int BLOCK = 256;
#ifndef ZZ_SWEEP_BLOCK
#define ZZ_SWEEP_BLOCK 512
#endif
// -DZZ_BOUNDS_CHECK (zigzag_b200/libzigzag_gpu_dbg.so, tests/test_gpu_debug_build.py): every index that is READ from memory before it
// addresses memory — the gene permutation, its inverse, slot lists, candidate slots, tile tickets, component numbers — is checked on
// the device; a violation prints its location and traps the kernel, which the next API call reports (compute-sanitizer is closed on
// the GPU pool this was developed on).
#ifdef ZZ_BOUNDS_CHECK
#define ZZ_DEV_ASSERT(cond) do { if (!(cond)) { printf("zigzag_gpu: device assert '%s' failed at %s:%d\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define ZZ_DEV_ASSERT(cond) ((void)0)
#endif
constexpr int SWEEP_BLOCK = ZZ_SWEEP_BLOCK;   // threads per CTA of the persistent sweep kernel
constexpr int RED_BLOCKS = 148 * 4;  // persistent-style reduction grid: 4 CTAs per SM

// ------------------------------------------------------------------------------------------------ kernels

constexpr int NSTAT_FIXED = 10;
constexpr int COMM_SMALL_NVEC = 4;   // gene sums exchanged by zz_run_generations: at most [2 parameter sets][2 values]
constexpr int MAX_NSTAT = 3 * (ZZ_MAX_COMP + 1) + NSTAT_FIXED;

// ---- sufficient-statistics all-reduce over NVLink peer memory, fused with the final reduction -------------------------------
// Every rank (one process per GPU) owns an inbox [2 slots][world][2 * nvec] of 64-bit words in its own HBM, mapped into every
// peer through CUDA IPC.  The vector travels in the "low-latency" format: each double is sent as two 8-byte words
// {low half | epoch << 32}, {high half | epoch << 32}.  An aligned 8-byte store is single-copy atomic, so a word whose upper half
// equals the current epoch carries valid data: no fence, no separate flag, ONE NVLink hop between "my vector is reduced" and
// "every peer can read it".  The CTA that holds the rank's reduced vector (1) stores its 2 * nvec words into
// inbox[slot][my_rank] of every peer (remote stores ride NVLink/NVSwitch), (2) polls its own inbox until the words of every peer
// show the epoch, (3) adds the `world` vectors in RANK ORDER (its own straight from shared memory), so every rank obtains the
// bit-identical sum (replicated hyper moves cannot diverge).  Slots alternate with the epoch; a rank cannot lap another by more
// than one epoch because it needs everybody's words first.  The poll is bounded (cm.timeout_ns of %globaltimer): on time-out the
// exchange is FAIL-STOP — bit 2 of the NaN/err flag is raised and every entry of `out` is NaN, so nothing downstream (the
// replicated scalar stage, the host's hyper moves) can consume words of an earlier epoch and let the ranks diverge silently.  (The first version — plain stores, __threadfence_system, a flag per peer, a
// second fence — cost 11-16 us per step at N = 2..8.)
struct CommArgs {
  int rank, world, nvec;                 // nvec = nval * C doubles per rank
  unsigned long long epoch;
  double *inbox[8];                      // inbox[r]: base of rank r's inbox (64-bit words), as mapped in THIS process
  unsigned long long *flags[8];          // unused since the low-latency format (kept: part of the exchanged IPC handle pair)
  int64_t base;                          // first word of the inbox region used by this exchange (statistics: 0; small gene sums: behind them)
  int stride;                            // words per (slot, rank) block of that region: fixed per region, so exchanges of different
                                         // lengths in the same region never overlap each other's slots
  long long timeout_ns;                  // bound of the poll (ZZ_COMM_TIMEOUT_S, default 20 s)
};
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// all threads of the CTA must call it; `vec` = the local vector in shared memory (complete: the caller has synchronised).
// One thread per (peer, word) sends; one thread per value polls its peers in RANK ORDER (the sum must be bit-identical on every
// rank) — but it first waits for ALL its words with the loads of the different peers in flight together, instead of paying one L2
// round trip per peer after the words have arrived.
__device__ __forceinline__ void allreduce_exchange(const double *vec, const CommArgs &cm, double *out, int *err_flag) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  const int slot = (int)(cm.epoch & 1ull), nw = 2 * cm.nvec;
  const unsigned ep = (unsigned)cm.epoch;
  for (int t = threadIdx.x; t < nw * cm.world; t += blockDim.x) {
    const int r = t / nw, w = t - r * nw;
    if (r == cm.rank) continue;
    const double v = vec[w >> 1];
    const unsigned half = (w & 1) ? (unsigned)__double2hiint(v) : (unsigned)__double2loint(v);
    const unsigned long long word = ((unsigned long long)ep << 32) | half;
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(cm.inbox[r]) + cm.base + ((int64_t)slot * cm.world + cm.rank) * cm.stride + w;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
  }
  const unsigned long long *in = reinterpret_cast<const unsigned long long *>(cm.inbox[cm.rank]) + cm.base + (int64_t)slot * cm.world * cm.stride;
  const unsigned long long t0 = globaltimer_ns();
  for (int idx = threadIdx.x; idx < cm.nvec; idx += blockDim.x) {
    unsigned long long a[8], b[8];
    bool all;
    do {                                   // the 2 (world - 1) loads of one round are independent: one round trip per round
      all = true;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r >= cm.world || r == cm.rank) continue;
        const volatile unsigned long long *p = in + (int64_t)r * cm.stride + 2 * idx;
        a[r] = p[0]; b[r] = p[1];
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r >= cm.world || r == cm.rank) continue;
        all = all && (unsigned)(a[r] >> 32) == ep && (unsigned)(b[r] >> 32) == ep;
      }
      if (!all && (long long)(globaltimer_ns() - t0) > cm.timeout_ns) { timed_out = 1; break; }
    } while (!all);
    double v = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r >= cm.world) continue;
      v += (r == cm.rank) ? vec[idx] : __hiloint2double((int)(unsigned)b[r], (int)(unsigned)a[r]);
    }
    out[idx] = v;
  }
  __syncthreads();
  if (timed_out) {                       // a peer never showed up: poison the result, raise the flag
    for (int idx = threadIdx.x; idx < cm.nvec; idx += blockDim.x) out[idx] = __longlong_as_double(0x7ff8000000000000LL);
    if (threadIdx.x == 0) atomicOr(err_flag, 4);
  }
}

// Fixed-order reduction of the per-CTA rows [C][nblk][nval] into `vec` (shared memory, nvec = nval*C doubles, followed by
// RED_THREADS doubles of scratch); whole CTA (>= RED_THREADS threads).  The rows were written by other CTAs: they are read
// through L2 (ld.cg).  Thread t owns statistic s = t % nval of row group j = t / nval, so the lanes of a warp read CONSECUTIVE
// doubles (one or two whole rows per load instruction).  Mapping lanes to rows instead (stride nval*8 bytes: 32 different cache
// lines per load) made the 384 loads of the final reduction cost 10 us per sweep.  The order is canonical — chain by chain,
// group j adds rows j, j + ng, ... in ascending order, then one thread per statistic adds the ng partial sums in group order,
// ng = RED_THREADS / nval — so every caller (sweep epilogue, k_reduce_final, k_reduce_allreduce) produces the same bits.
constexpr int RED_THREADS = 256;
static_assert(SWEEP_BLOCK >= RED_THREADS && BLOCK >= RED_THREADS, "reduce_rows_to_smem needs RED_THREADS threads");
__device__ __forceinline__ void reduce_rows_to_smem(const double *partials, int nblk, int nval, int nvec, double *vec) {
  double *scratch = vec + nvec;
  const int C = nvec / nval, ng = RED_THREADS / nval;
  const int t = threadIdx.x, s = t % nval, j = t / nval;
  for (int y = 0; y < C; ++y) {
    if (j < ng) {
      const double *base = partials + (int64_t)y * nblk * nval + s;
      double v = 0;
      for (int r0 = j; r0 < nblk; r0 += ng * 8) {
        double tt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int r = r0 + u * ng; tt[u] = r < nblk ? __ldcg(base + (int64_t)r * nval) : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) v += tt[u];
      }
      scratch[j * nval + s] = v;
    }
    __syncthreads();
    if (t < nval) {
      double a = 0;
      for (int k = 0; k < ng; ++k) a += scratch[k * nval + t];
      vec[y * nval + t] = a;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(BLOCK) k_reduce_allreduce(const double *partials, int nblk, int nval, int C, double *out, CommArgs cm,
                                                             int *err_flag) {
  __shared__ double vec[MAX_NSTAT * 8 + RED_THREADS];   // local reduced vector (nvec <= MAX_NSTAT * 8 is checked by the host) + scratch
  (void)C;
  reduce_rows_to_smem(partials, nblk, nval, cm.nvec, vec);
  allreduce_exchange(vec, cm, out, err_flag);
}


struct SweepArgs {
  int64_t G, g_begin, g_end, gene_offset;
  int C, R, K, tune;
  uint32_t moves;
  uint64_t seed, step;
  const double *Xg, *gl;
  double *Yg, *vg, *xglik, *vgprob;
  const double *ty, *tv;
  uint16_t *flags;
  ulonglong2 *win_y, *win_v;
  const zz_hyper *hyper;
  const int32_t *perm;              // device slot -> original (caller's) local gene index; Philox is keyed by the original index
  const double *draws[9];
  uint8_t *dec[4];
  int *nan_flag;
};

__global__ void __launch_bounds__(BLOCK) k_sweep(const SweepArgs a) {
  __shared__ HyperS H;
  const int chain = blockIdx.y;
  load_hyper(H, a.hyper + chain);
  const int64_t g = a.g_begin + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= a.g_end) return;
  const int64_t i = (int64_t)chain * a.G + g;
  const uint64_t gid = (uint64_t)(a.gene_offset + a.perm[g]);
  Gene gn;
  gn.y = a.Yg[i]; gn.v = a.vg[i]; gn.xglik = a.xglik[i]; gn.vgprob = a.vgprob[i]; gn.f = a.flags[i];
  const double gl = a.gl[g];
  const double *xg = a.Xg + g;
  bool saw_nan = false;
  uint32_t b0[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};
  if (!a.draws[0] && (a.moves & (ZZ_MOVE_YG | ZZ_MOVE_VARIANCE_G))) philox_block(a.seed, chain, a.step, 0, gid, b0);
  if (!a.draws[3] && (a.moves & (ZZ_MOVE_VARIANCE_G | ZZ_MOVE_ALLOC | ZZ_MOVE_ALLOC_WITHIN))) philox_block(a.seed, chain, a.step, 1, gid, b1);

  if (a.moves & ZZ_MOVE_YG) {
    double z, u;
    if (a.draws[0]) { z = a.draws[0][i]; u = a.draws[1][i]; }
    else { double z1; box_muller(u32_open(b0[0]), u32_open(b0[1]), z, z1); u = u32_open(b0[2]); }
    const bool proposed = !(gn.f & F_INSPIKE);
    double ratio;
    const int acc = move_yg(H, gn, xg, a.G, gl, a.ty[i], z, u, ratio);
    saw_nan |= proposed && isnan(ratio);
    if (a.tune && proposed) { ulonglong2 w = a.win_y[i]; window_push(w, acc); a.win_y[i] = w; }  // Pr:312-322
    if (a.dec[0]) a.dec[0][i] = (uint8_t)acc;
  }
  if (a.moves & ZZ_MOVE_VARIANCE_G) {
    double u1, u2;
    if (a.draws[2]) { u1 = a.draws[2][i]; u2 = a.draws[3][i]; }
    else { u1 = u32_open(b0[3]); u2 = u32_open(b1[0]); }
    const bool proposed = !(gn.f & F_INSPIKE);
    double ratio;
    const int acc = move_variance_g(H, gn, xg, a.G, gl, a.tv[i], u1, u2, ratio);
    saw_nan |= proposed && isnan(ratio);
    if (a.tune && proposed) { ulonglong2 w = a.win_v[i]; window_push(w, acc); a.win_v[i] = w; }  // Pr:40-50
    if (a.dec[1]) a.dec[1][i] = (uint8_t)acc;
  }
  if (a.moves & ZZ_MOVE_ALLOC) {
    double u_ai, u_w;
    if (a.draws[4]) { u_ai = a.draws[4][i]; u_w = a.draws[5][i]; }
    else { u_ai = u32_open(b1[1]); u_w = u32_open(b1[2]); }
    double pa;
    const int word = move_alloc(H, gn, u_ai, u_w, pa);
    if (a.dec[2]) a.dec[2][i] = (uint8_t)word;
  } else if (a.moves & ZZ_MOVE_ALLOC_WITHIN) {
    double u_w;
    if (a.draws[5]) u_w = a.draws[5][i];
    else u_w = u32_open(b1[2]);
    if (gn.f & F_ACTIVE) { const int kn = within_active(H, gn.y, u_w); if (kn > 0) gn.f = flag_set_k(gn.f, kn); }
    if (a.dec[2]) a.dec[2][i] = (uint8_t)((gn.f & F_ACTIVE) ? flag_k(gn.f) : 0);
  }
  if (a.moves & ZZ_MOVE_SPIKE) {
    double zy, zv, u;
    if (a.draws[6]) { zy = a.draws[6][i]; zv = a.draws[7][i]; u = a.draws[8][i]; }
    else {
      uint32_t b2[4];
      philox_block(a.seed, chain, a.step, 2, gid, b2);
      box_muller(u32_open(b2[0]), u32_open(b2[1]), zy, zv); u = u32_open(b2[2]);
    }
    const bool proposed = (gn.f & F_ALLZERO) && !(gn.f & F_ACTIVE);
    double ratio;
    const int acc = move_spike(H, gn, xg, a.G, gl, zy, zv, u, ratio);
    saw_nan |= proposed && isnan(ratio);
    if (a.dec[3]) a.dec[3][i] = (uint8_t)acc;
  }
  a.Yg[i] = gn.y; a.vg[i] = gn.v; a.xglik[i] = gn.xglik; a.vgprob[i] = gn.vgprob; a.flags[i] = (uint16_t)gn.f;
  if (saw_nan) atomicOr(a.nan_flag, 1);
}

// zz_sweep_host: per-gene state entering / leaving in the CALLER's gene order, either through staging planes in device memory
// (filled / drained by cudaMemcpyAsync) or — when the caller's arrays are pinned, device-accessible host memory — directly.
struct HostChunk {
  int64_t c0, n;
  const int32_t *inv;                                   // caller gene -> device slot
  const double *sYg, *svg, *sty, *stv; const int32_t *sai, *sw, *ss;    // inbound planes  [G], caller order
  double *dYg, *dvg, *dxl, *dvp; int32_t *dai, *dw, *ds;                // outbound planes [G], caller order
  double *ty, *tv;                                      // the handle's tuning-parameter arrays (read-only in SweepArgs)
};

// ---- production sweep: same moves and draws as k_sweep in the reorganised arithmetic of zz_fast.cuh -------------------
struct FastArgs {
  SweepArgs s;
  double *lps;                      // [C][G] A(y): sum of the detection log-terms at the current Yg
  const double *exp_tab; const double2 *log_tab;
  const uint64_t *nd_mask; const double *xbar, *xss;  // per-gene data statistics (k_prepare_data)
  const FastH *fasth;               // [C] derived hyper-parameters (k_prepare_hyper)
  double *partials;                 // [C][rows][nstat] per-block sufficient statistics (rows = main blocks + spike blocks)
  int64_t rows_per_chain;
  int want_stats, do_spike_move;
  const int32_t *slot_list;         // list mode (zz_sweep_host chunks): position p in [g_begin, g_end) addresses gene slot slot_list[p]
  const int *gate; int gate_stride; // k_refresh_lps in the device-resident loop: chain c is rebuilt only if gate[c * gate_stride] != 0 (accepted mhP_x)
  int host_io; HostChunk hio;       // list mode + host_io: the tile itself reads its genes' state from hio.s* and writes hio.d* (caller order)
  // fused epilogue: the last CTA to finish reduces the rows (and all-reduces them over peer memory when asked) into stats_out
  unsigned *done_counter; double *stats_out; int do_allreduce; CommArgs cm;
};

__global__ void k_prepare_hyper(int C, const zz_hyper *hyper, FastH *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) derive_fast(out[c], hyper + c);
}

// Xg [R][G] -> per-gene (nd_mask, n_det via popcount, xbar, ss); run once per zz_upload_data
__global__ void __launch_bounds__(BLOCK) k_prepare_data(int64_t G, int R, const double *__restrict__ Xg, uint64_t *nd_mask, double *xbar, double *xss) {
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= G) return;
  uint64_t m = 0; double s = 0; int n = 0;
  for (int r = 0; r < R; ++r) {
    const double x = Xg[(int64_t)r * G + g];
    if (x == -INFINITY) m |= (1ull << r); else { s += x; n++; }
  }
  const double mean = n > 0 ? s / n : 0.0;
  double ss = 0, c = 0;
  for (int r = 0; r < R; ++r) {
    const double x = Xg[(int64_t)r * G + g];
    if (x != -INFINITY) { const double d = x - mean; ss = fma(d, d, ss); c += d; }
  }
  // `mean` is rounded: sum (x - y)^2 = ss + 2 (mean - y) c + n (mean - y)^2 with c = sum (x - mean) ~ 1e-16; fold the exact
  // correction of the mean so that the cross term vanishes to second order
  const double mean2 = n > 0 ? mean + c / n : 0.0;
  ss = n > 0 ? ss - c * c / n : 0.0;
  nd_mask[g] = m; xbar[g] = mean2; xss[g] = ss < 0 ? 0.0 : ss;
}

// Persistent kernel: blockIdx.y = chain; warp w of the chain's grid handles the 32-gene tiles w, w + W, w + 2W, ...
// One table/hyper load and one barrier per block for the whole sweep; warps never wait for each other afterwards.
// PROD = true is the production instantiation (Philox draws, no decision export); PROD = false additionally serves the
// "identical uniforms" test mode (explicit draws) and exports decisions.
// mhSpikeAllocation (Pr:528-638) for one all-zero, inactive gene; its likelihood needs no Xg (every library is -Inf).
// Works on the gene's global-memory state (after the main moves stored it) and returns what the statistics need.  Out of
// line: only the warps that own all-zero genes (the last slots of the sorted gene order) ever call it.
struct SpikeOut { double y, lv, xglik, vgprob; uint32_t f; int acc; bool finite_state; };

template <int RT>
__device__ __noinline__ void spike_move(const FastH &H, const MathTabs &T, const FastArgs &fa, int chain, int64_t g, int64_t i, SpikeOut &o) {
  const SweepArgs &a = fa.s;
  uint32_t f = a.flags[i];
  double y = a.Yg[i], v = a.vg[i], xglik = a.xglik[i], vgprob = a.vgprob[i];
  int acc_s = 0;
  if (!(f & F_ACTIVE)) {                                              // Pr:533: all-zero AND inactive
    const bool cur_in = (f & F_INSPIKE) != 0, prop_in = !cur_in;      // Pr:535
    double zy, zv, u;
    if (a.draws[6]) { zy = a.draws[6][i]; zv = a.draws[7][i]; u = a.draws[8][i]; }
    else {
      uint32_t b2[4];
      philox_block(a.seed, chain, a.step, 2, (uint64_t)(a.gene_offset + a.perm[g]), b2);
      box_muller(u32_open(b2[0]), u32_open(b2[1]), zy, zv); u = u32_open(b2[2]);
    }
    const double yp = fma(H.sd_i, zy, H.im);                          // Pr:545
    const double lsp = fma(H.s1, yp, H.s0);                           // log(Sg_proposed), Pr:547
    const double lvp = (lsp + H.tau) + H.rtau * zv;                   // Pr:549: variance_g' = exp(.)
    const double vp = fexp(lvp, T.exp_t);
    double PVp = vg_prior_from_q(H, lvp, vg_q(H, lvp, yp));           // Pr:556
    const double lv = flog(v, T.log_t);
    double PVc = vg_prior_from_q(H, lv, vg_q(H, lv, y));              // Pr:558
    if (H.vub != INFINITY) { PVp -= plnorm_term(H, yp); PVc -= plnorm_term(H, y); }
    const double dp = yp - H.im, dc = y - H.im;
    const double pygp = prop_in ? H.log_sp : H.log_1msp - H.log_norm_i - dp * dp * H.inv2iv;   // Pr:560
    const double pyg = cur_in ? H.log_sp : H.log_1msp - H.log_norm_i - dc * dc * H.inv2iv;     // Pr:563
    const double pp = prop_in ? H.log_sp : (pygp + PVp);              // Pr:569-570
    const double pc = cur_in ? H.log_sp : (pyg + PVc);                // Pr:572-573
    double Ap = 0;
    if (!prop_in) Ap = detect_logsum<RT>(H, T, ~0ull, a.gl[g] * fexp(yp, T.exp_t));
    double Lp = prop_in ? 0.0 : times_beta(H, Ap);                    // Pr:579 (P:117 gives log(1*1) = 0 in the spike)
    double Lc = xglik;                                                // Pr:581
    if (Lp == -INFINITY && Lc == -INFINITY) { Lp = 0; Lc = 0; }       // Pr:583-585
    const double HR = prop_in ? (pyg - H.log_1msp + PVc) : -(pygp - H.log_1msp + PVp);   // Pr:592-595
    double ratio = H.temp * (Lp - Lc + pp - pc) + HR;                 // Pr:601
    if (vp > H.vub) ratio = -INFINITY;                                // Pr:603
    acc_s = (flog_core(u, T.log_t) < ratio);                          // Pr:605
    if (isnan(ratio)) atomicOr(a.nan_flag, 1);
    if (acc_s) f ^= F_INSPIKE;                                        // Pr:616
    a.flags[i] = (uint16_t)f;
    if (!prop_in) {                                                   // [proposed_out_spike_idx], Pr:620-636
      if (acc_s) { y = yp; v = vp; a.Yg[i] = yp; a.vg[i] = vp; fa.lps[i] = Ap; }
      vgprob = acc_s ? PVp : PVc;
      a.vgprob[i] = vgprob;
    }
    xglik = acc_s ? Lp : Lc;                                          // Pr:630-632
    a.xglik[i] = xglik;
  }
  o.y = y; o.lv = (f & F_INSPIKE) ? 0.0 : flog(v, T.log_t); o.xglik = xglik; o.vgprob = vgprob; o.f = f; o.acc = acc_s;
  o.finite_state = isfinite(y) && isfinite(v);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// cp.async of one lane's ten input words of gene slot g (chain `chain`) into its staging column `stg` ([slot * 32])
__device__ __forceinline__ void stage_tile(const SweepArgs &a, const FastArgs &fa, int chain, int64_t g, double *stg) {
  const int64_t i = (int64_t)chain * a.G + g;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stg);
  const void *src[9] = {a.Yg + i, a.vg + i, fa.lps + i, a.ty + i, a.tv + i, a.gl + g, fa.xbar + g, fa.xss + g, fa.nd_mask + g};
#pragma unroll
  for (int j = 0; j < 9; ++j) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + j * 256), "l"(src[j]) : "memory");
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 9 * 256), "l"(a.perm + g) : "memory");
}

// the same with an L2 evict-first policy: the per-gene arrays of a sweep are touched once per generation and are larger than L2
// together, so they should not displace what IS reused between generations (the scalar stage's code and state, the constants)
__device__ __forceinline__ uint64_t l2_evict_first_policy() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void stage_tile_ef(const SweepArgs &a, const FastArgs &fa, int chain, int64_t g, double *stg, uint64_t pol) {
  const int64_t i = (int64_t)chain * a.G + g;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stg);
  const void *src[9] = {a.Yg + i, a.vg + i, fa.lps + i, a.ty + i, a.tv + i, a.gl + g, fa.xbar + g, fa.xss + g, fa.nd_mask + g};
#pragma unroll
  for (int j = 0; j < 9; ++j) asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dst + j * 256), "l"(src[j]), "l"(pol) : "memory");
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(dst + 9 * 256), "l"(a.perm + g), "l"(pol) : "memory");
}

template <int RT>
__device__ __noinline__ double detect_logsum_slow(const FastH &H, const MathTabs &T, uint64_t nd_mask, double t) { return detect_logsum<RT>(H, T, nd_mask, t); }

// KT > 0 selects the straight-line production tile (all four moves, Philox draws, variance_g_upper_bound = Inf, K = KT active
// components): every lane executes the same branch-free instruction stream (decisions are applied with selects), which lets
// the scheduler overlap the independent exp/log/reciprocal chains of the three moves instead of serialising them.
template <int RT, bool PROD, int KT>
__global__ void __launch_bounds__(SWEEP_BLOCK, ZZ_SWEEP_MINB) k_sweep_fast(const FastArgs fa) {
  __shared__ FastH H;
  __shared__ MathTabs T;
  __shared__ double red[SWEEP_BLOCK / 32];
  extern __shared__ double acc[];                     // [nstat][BLOCK] when want_stats
  const SweepArgs &a = fa.s;
  const int chain = blockIdx.y, lane = threadIdx.x & 31;
  const int nstat = 3 * (a.K + 1) + NSTAT_FIXED_;
  asm volatile("griddepcontrol.launch_dependents;");   // see launch_pdl
  if (fa.want_stats) for (int s = 0; s < nstat; ++s) acc[s * SWEEP_BLOCK + threadIdx.x] = 0.0;
  load_tabs(T, fa.exp_tab, fa.log_tab);                // constants: may be read while the previous kernel still runs
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads state the previous kernel may have written
  load_hyper_fast(H, fa.fasth + chain);
  const int64_t n = a.g_end - a.g_begin;
  const int64_t ntiles = (n + 31) / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * (SWEEP_BLOCK / 32) + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * (SWEEP_BLOCK / 32);
  bool saw_nan = false;
  if constexpr (KT > 0) {
    // per-warp staging slots [10][32] doubles behind the statistic columns: Yg, variance_g, lps, tuning parameters, gene length,
    // xbar, ss, nd_mask, perm of the warp's NEXT tile arrive by cp.async while the current tile computes
    double *stg = acc + (int64_t)nstat * SWEEP_BLOCK + (threadIdx.x >> 5) * 320 + lane;
    uint32_t f_next = 0;
    {
      const int64_t g0 = a.g_begin + (ntiles - 1 - warp0) * 32 + lane;
      if (warp0 < ntiles && g0 < a.g_end) { stage_tile(a, fa, chain, g0, stg); f_next = a.flags[(int64_t)chain * a.G + g0]; }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int64_t t_ = warp0; t_ < ntiles; t_ += nwarps) {
      const int64_t tile = ntiles - 1 - t_;
      const int64_t g = a.g_begin + tile * 32 + lane;
      const int64_t i = (int64_t)chain * a.G + g;
      if (g >= a.g_end) { asm volatile("cp.async.wait_group 0;" ::: "memory"); continue; }
      // ---------------- straight-line production tile ----------------
      asm volatile("cp.async.wait_group 0;" ::: "memory");            // this lane's copies of the current tile have landed
      const double y0 = stg[0 * 32], v0 = stg[1 * 32], A0 = stg[2 * 32], ty = stg[3 * 32], tv = stg[4 * 32], gl = stg[5 * 32];
      const double xbar = stg[6 * 32], ss = stg[7 * 32];
      const uint64_t mask = (uint64_t)__double_as_longlong(stg[8 * 32]);
      const uint64_t gid = (uint64_t)(a.gene_offset + (int64_t)(uint32_t)__double2loint(stg[9 * 32]));
      uint32_t f = f_next;
      const int64_t t_n = t_ + nwarps;
      const int64_t g_n = a.g_begin + (ntiles - 1 - t_n) * 32 + lane;
      const bool next_valid = t_n < ntiles && g_n < a.g_end;
      if (next_valid) f_next = a.flags[(int64_t)chain * a.G + g_n];     // plain load, consumed one tile later
      const double nd_d = (double)((RT > 0 ? RT : H.R) - __popcll(mask));
      const bool live = !(f & F_INSPIKE);
      uint32_t b0[4], b1[4];
      philox_block(a.seed, chain, a.step, 0, gid, b0);
      philox_block(a.seed, chain, a.step, 1, gid, b1);
      const double u_y = u32_to_open01(b0[2]), u1 = u32_to_open01(b0[3]), u2 = u32_to_open01(b1[0]);
      const double u_ai = u32_to_open01(b1[1]), u_w = u32_to_open01(b1[2]);
      const double z = fsqrt_nb(-2.0 * flog_nb(u32_to_open01(b0[0]), T.log_t)) * cospi(2.0 * u32_to_open01(b0[1]));
      const double lv = flog_nb(v0, T.log_t), rv = frcp_nb(v0);
      // every staged word of this tile has been consumed (z, lv depend on them): refill the slots with the warp's next tile, whose
      // HBM latency then hides behind this tile's ~1600 instructions
      if (next_valid) stage_tile(a, fa, chain, g_n, stg);
      asm volatile("cp.async.commit_group;" ::: "memory");
      bool bad = !is_pos_normal(v0);                                      // poisoned state: reported through the NaN flag
      // ---- mhYg (Pr:250-333)
      const double yp = y0 + ty * z;                                     // Pr:257
      const double tp = gl * fexp_nb(yp, T.exp_t);                       // U:157
      const bool saturated = __all_sync(__activemask(), !live || tp * H.alpha_min >= 37.5);
      // (uclamp_ok is checked by the host when it picks this kernel; in the device-resident loop alpha_r moves on the device, so the
      // per-library-clamp version stays reachable, out of line)
      const double Ap = saturated ? (mask ? -INFINITY : 0.0)
                        : H.uclamp_ok ? detect_logsum_sl<(RT > 0 && RT <= 18) ? RT : 1>(H, T, mask, tp) : detect_logsum_slow<RT>(H, T, mask, tp);
      const double t0 = xbar - y0, t1 = xbar - yp;
      const double D0 = fma(nd_d * t0, t0, ss), Dp = fma(nd_d * t1, t1, ss);
      const double dLX = times_beta(H, (Ap - A0) - 0.5 * (Dp - D0) * rv);
      const double q0 = vg_q(H, lv, y0), qp = vg_q(H, lv, yp);
      const double dPV = -0.5 * (qp * qp - q0 * q0);
      const bool act0 = (f & F_ACTIVE) != 0;
      const int k0 = flag_k(f) - 1;
      const double m2 = act0 ? H.am[k0] : H.im, i2 = act0 ? H.inv2av[k0] : H.inv2iv;
      const double e0 = y0 - m2, e1 = yp - m2;
      const double dL2 = -(e1 * e1 * i2 - e0 * e0 * i2);
      const double ratio_y = H.temp * (dLX + dL2 + dPV);                 // Pr:296
      const bool acc_y = live && (flog_nb(u_y, T.log_t) < ratio_y);      // Pr:302
      const double y1 = acc_y ? yp : y0, A1 = acc_y ? Ap : A0, D1 = acc_y ? Dp : D0, q1 = acc_y ? qp : q0;
      // ---- mhVariance_g (Pr:7-56)
      const double arg = tv * (u1 - 0.5);                                // MP:5
      const double sf = fexp_nb(arg, T.exp_t), vp = v0 * sf, rvp = rv * fexp_nb(-arg, T.exp_t);
      const double dB = -nd_d * (0.5 * arg) - 0.5 * D1 * (rvp - rv);
      const double qv = q1 + arg * H.irt;
      double ratio_v = H.temp * (times_beta(H, dB) + (-arg - 0.5 * (qv * qv - q1 * q1))) + arg;   // Pr:24
      if (!(A1 > -INFINITY) && H.beta != 0.0) ratio_v = NAN;             // -Inf current likelihood: as the reference, NaN => reject
      const bool acc_v = live && (flog_nb(u2, T.log_t) < ratio_v);       // Pr:35
      const double v1 = acc_v ? vp : v0, lv1 = acc_v ? lv + arg : lv, rv1 = acc_v ? rvp : rv;
      // ---- gibbsAllocationActiveInactive + WithinActive (Pr:335-410)
      const double dI = y1 - H.im;
      const double L0 = live ? H.c0 * fexp_nb(-(dI * dI) * H.inv2iv, T.exp_t) : H.sp;
      double Lk[KT], L1 = 0;
#pragma unroll
      for (int k = 0; k < KT; ++k) { const double d = y1 - H.am[k]; Lk[k] = H.ck[k] * fexp_nb(-(d * d) * H.inv2av[k], T.exp_t); L1 += Lk[k]; }
      const double num = live ? H.wa * L1 : 0.0, den = H.one_m_wa * L0 + num;
      double pa = num * frcp_nb(den);                                    // Pr:348
      if (!is_pos_normal(den)) { pa = num / den; if (isnan(pa)) pa = H.wa; }   // Pr:349 (rare)
      const bool a_new = (f & F_PINNED) || (u_ai < pa);                  // Pr:353-355
      const double thr = u_w * L1;
      double cum = 0; int cnt = 0;
#pragma unroll
      for (int k = 0; k < KT; ++k) { cum += Lk[k]; cnt += (cum < thr); } // Pr:398-402
      const int kn = cnt + 1 > KT ? KT : cnt + 1;
      const bool row_ok = L1 > 0 && L1 < INFINITY;                       // NaN rows keep the old k, Pr:403
      f = (f & ~F_ACTIVE) | (a_new ? F_ACTIVE : 0u);
      if (a_new && row_ok) f = flag_set_k(f, kn);
      // ---- caches, stores
      double xglik = times_beta(H, A1 + (-nd_d * (LN_SQRT_2PI + 0.5 * lv1) - 0.5 * D1 * rv1));
      double vgprob = vg_prior_from_q(H, lv1, vg_q(H, lv1, y1));
      a.flags[i] = (uint16_t)f;
      if (live) { a.Yg[i] = y1; a.vg[i] = v1; fa.lps[i] = A1; a.xglik[i] = xglik; a.vgprob[i] = vgprob; }
      else {
        xglik = times_beta(H, (f & F_ALLZERO) ? 0.0 : -INFINITY);        // P:117
        if (fa.want_stats && (f & F_ALLZERO)) xglik = a.xglik[i];        // the cache as the last spike move left it
      }
      if (a.tune && live) {                                              // Pr:312-322, Pr:40-50
        ulonglong2 w = a.win_y[i]; window_push(w, acc_y); a.win_y[i] = w;
        w = a.win_v[i]; window_push(w, acc_v); a.win_v[i] = w;
      }
      saw_nan |= live && (isnan(ratio_y) || isnan(ratio_v) || bad);
      double ys = y1, lvs = lv1;
      bool finite_state = isfinite(y1) && isfinite(v1);
      if ((f & F_ALLZERO) && fa.do_spike_move) {
        SpikeOut so;
        spike_move<RT>(H, T, fa, chain, g, i, so);
        ys = so.y; lvs = so.lv; f = so.f; xglik = so.xglik; vgprob = so.vgprob; finite_state = so.finite_state;
      }
      if (fa.want_stats) stats_add<SWEEP_BLOCK>(acc, a.K, true, ys, lvs, f, xglik, vgprob, finite_state);
    }
  } else {
  // Tiles are taken from the END of the sorted gene order first: the all-zero genes (extra spike move) and the weakly
  // expressed genes (full detection loop) run while every warp is busy; the last, partially filled round consists of
  // highly expressed tiles whose detection term is saturated, which keeps the tail short.
  for (int64_t t_ = warp0; t_ < ntiles; t_ += nwarps) {
    const int64_t tile = ntiles - 1 - t_;
    int64_t g = a.g_begin + tile * 32 + lane;
#ifndef ZZ_NO_PREFETCH
    {   // pull the next tile's inputs towards L1 while this tile computes (no registers, no shared memory)
      const int64_t gn_ = g - nwarps * 32;
      if (!fa.slot_list && gn_ >= a.g_begin && gn_ < a.g_end) {
        const int64_t in_ = (int64_t)chain * a.G + gn_;
        prefetch_l1(a.Yg + in_); prefetch_l1(a.vg + in_); prefetch_l1(fa.lps + in_); prefetch_l1(a.ty + in_); prefetch_l1(a.tv + in_);
        prefetch_l1(fa.nd_mask + gn_); prefetch_l1(fa.xbar + gn_); prefetch_l1(fa.xss + gn_); prefetch_l1(a.gl + gn_);
        if ((lane & 3) == 0) prefetch_l1(a.flags + in_);
      }
    }
#endif
    if (g >= a.g_end) continue;
    const int64_t pos = g;
    if (fa.slot_list) g = fa.slot_list[g];            // a chunk of genes in the caller's order: scattered slots
    ZZ_DEV_ASSERT(g >= 0 && g < a.G);
    const int64_t i = (int64_t)chain * a.G + g;
    if (fa.host_io) {
      // The gene's state comes straight from the caller's (pinned, device-mapped) arrays: PCIe reads issued by this warp while
      // other resident warps compute and write back (a cp.async double buffer of the next tile's words was measured: no gain,
      // the PCIe link is the bound).  The same work as k_host_in follows, fused into the tile.
      const HostChunk &hc = fa.hio;
      const double y = hc.sYg[pos], v_in = hc.svg[pos], ty_in = hc.sty[pos], tv_in = hc.stv[pos];
      const int ai_in = hc.sai[pos], w_in = hc.sw[pos], s_in = hc.ss[pos];
      a.Yg[i] = y; a.vg[i] = v_in; hc.ty[i] = ty_in; hc.tv[i] = tv_in;
      uint32_t f = a.flags[i];
      f = (f & ~F_ACTIVE) | (ai_in ? F_ACTIVE : 0u);
      f = flag_set_k(f, w_in < 1 ? 1 : w_in);
      f = (f & ~F_INSPIKE) | (s_in ? F_INSPIKE : 0u);
      a.flags[i] = (uint16_t)f;
      const double A = (f & F_INSPIKE) ? 0.0 : detect_logsum<RT>(H, T, fa.nd_mask[g], a.gl[g] * fexp(y, T.exp_t));
      fa.lps[i] = A;
      if (f & F_ALLZERO) a.xglik[i] = (f & F_INSPIKE) ? 0.0 : times_beta(H, A);
    }
    const uint64_t gid = (uint64_t)(a.gene_offset + a.perm[g]);
    FastGene gn;
    gn.y = a.Yg[i]; gn.v = a.vg[i]; gn.A = fa.lps[i]; gn.f = a.flags[i];
    GeneData gd;
    gd.nd_mask = fa.nd_mask[g]; gd.xbar = fa.xbar[g]; gd.ss = fa.xss[g];
    gd.ndet = (RT > 0 ? RT : H.R) - __popcll(gd.nd_mask);
    const double gl = a.gl[g];
    const bool live = !(gn.f & F_INSPIKE);            // mhYg / mhVariance_g only touch out_spike_idx (Pr:11, Pr:252-257)
    double xglik = 0, vgprob = 0;
    gn.lv = 0; gn.rv = 0;
    uint32_t b0[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};
    const bool own_draws = PROD || !a.draws[0];
    if (own_draws && (a.moves & (ZZ_MOVE_YG | ZZ_MOVE_VARIANCE_G))) philox_block(a.seed, chain, a.step, 0, gid, b0);
    if (own_draws && (a.moves & (ZZ_MOVE_VARIANCE_G | ZZ_MOVE_ALLOC))) philox_block(a.seed, chain, a.step, 1, gid, b1);
    if (live) { gn.lv = flog(gn.v, T.log_t); gn.rv = 1.0 / gn.v; }
    if (a.moves & ZZ_MOVE_YG) {
      int acc_y = 0;
      double z, u, ratio;
      if (own_draws) { z = normal_cos(u32_to_open01(b0[0]), u32_to_open01(b0[1]), T); u = u32_to_open01(b0[2]); }
      else { z = a.draws[0][i]; u = a.draws[1][i]; }
      const double yp = gn.y + (live ? a.ty[i] : 0.0) * z;               // Pr:257
      const double tp = gl * fexp(yp, T.exp_t);                         // U:157
      const bool saturated = __all_sync(__activemask(), !live || tp * H.alpha_min >= 37.5);
      if (live) {
        acc_y = fast_yg<RT>(H, T, gn, gd, yp, tp, saturated, u, ratio);
        saw_nan |= isnan(ratio);
        if (a.tune) { ulonglong2 w = a.win_y[i]; window_push(w, acc_y); a.win_y[i] = w; }   // Pr:312-322
      }
      if (!PROD && a.dec[0]) a.dec[0][i] = (uint8_t)acc_y;
    }
    if (a.moves & ZZ_MOVE_VARIANCE_G) {
      int acc_v = 0;
      if (live) {
        double u1, u2, ratio;
        if (own_draws) { u1 = u32_to_open01(b0[3]); u2 = u32_to_open01(b1[0]); }
        else { u1 = a.draws[2][i]; u2 = a.draws[3][i]; }
        acc_v = fast_vg(H, T, gn, gd, a.tv[i], u1, u2, ratio);
        saw_nan |= isnan(ratio);
        if (a.tune) { ulonglong2 w = a.win_v[i]; window_push(w, acc_v); a.win_v[i] = w; }   // Pr:40-50
      }
      if (!PROD && a.dec[1]) a.dec[1][i] = (uint8_t)acc_v;
    }
    if (a.moves & ZZ_MOVE_ALLOC) {
      double u_ai, u_w;
      if (own_draws) { u_ai = u32_to_open01(b1[1]); u_w = u32_to_open01(b1[2]); }
      else { u_ai = a.draws[4][i]; u_w = a.draws[5][i]; }
      const int word = fast_alloc(H, T, gn.y, gn.f, u_ai, u_w);
      if (!PROD && a.dec[2]) a.dec[2][i] = (uint8_t)word;
      a.flags[i] = (uint16_t)gn.f;
    }
    if (live) {
      a.Yg[i] = gn.y; a.vg[i] = gn.v; fa.lps[i] = gn.A;
      // the reference's caches, re-expressed: XgLikelihood = beta (A + B), variance_g_probability = prior(v | Sg(y))
      xglik = xg_lik_fast(H, gn, gd);
      vgprob = vg_prior_from_q(H, gn.lv, vg_q(H, gn.lv, gn.y));
      if (H.vub != INFINITY) vgprob -= plnorm_term(H, gn.y);
      a.xglik[i] = xglik; a.vgprob[i] = vgprob;
    } else {
      xglik = times_beta(H, (gn.f & F_ALLZERO) ? 0.0 : -INFINITY);       // P:117
      if (fa.want_stats && (gn.f & F_ALLZERO)) xglik = a.xglik[i];       // the cache as the last spike move left it (Pr:583-585, 630-632)
    }
    bool finite_state = isfinite(gn.y) && isfinite(gn.v);
    if ((gn.f & F_ALLZERO) && fa.do_spike_move) {     // the last slots of the sorted gene order: whole warps take this branch
      SpikeOut so;
      spike_move<RT>(H, T, fa, chain, g, i, so);
      gn.y = so.y; gn.lv = so.lv; gn.f = so.f; xglik = so.xglik; vgprob = so.vgprob; finite_state = so.finite_state;
      if (!PROD && a.dec[3]) a.dec[3][i] = (uint8_t)so.acc;
    }
    if (fa.want_stats) stats_add<SWEEP_BLOCK>(acc, a.K, true, gn.y, gn.lv, gn.f, xglik, vgprob, finite_state);
    if (fa.host_io) {                                 // k_host_out, fused: the updated state as the moves (and the spike move) left it
      const HostChunk &hc = fa.hio;
      hc.dYg[pos] = a.Yg[i]; hc.dvg[pos] = a.vg[i]; hc.dxl[pos] = a.xglik[i]; hc.dvp[pos] = a.vgprob[i];
      const uint32_t fo = a.flags[i];
      hc.dai[pos] = (fo & F_ACTIVE) ? 1 : 0; hc.dw[pos] = flag_k(fo); hc.ds[pos] = (fo & F_INSPIKE) ? 1 : 0;
    }
  }
  }
  if (saw_nan) atomicOr(a.nan_flag, 1);
  if (fa.want_stats) {
    __syncthreads();
    stats_flush<SWEEP_BLOCK>(acc, nstat, fa.partials + ((int64_t)chain * fa.rows_per_chain + blockIdx.x) * nstat, red);
    if (fa.stats_out) {   // "last CTA done" epilogue: no second launch for the final reduction / the collective
      __shared__ unsigned ticket;
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); ticket = atomicAdd(fa.done_counter, 1u); }
      __syncthreads();
      if (ticket == gridDim.x * gridDim.y - 1) {
        __threadfence();
        double *vec = acc;                         // the statistic columns are dead now: reuse them as the reduced vector
        const int nvec = nstat * (int)gridDim.y;
        reduce_rows_to_smem(fa.partials, (int)fa.rows_per_chain, nstat, nvec, vec);
        if (fa.do_allreduce) allreduce_exchange(vec, fa.cm, fa.stats_out, a.nan_flag);
        else for (int idx = threadIdx.x; idx < nvec; idx += blockDim.x) fa.stats_out[idx] = vec[idx];
        if (threadIdx.x == 0) *fa.done_counter = 0u;
      }
    }
  }
}

// lps <- A(Yg) for every out-of-spike gene (after anything that changed Yg or alpha_r outside the fast sweep)
template <int RT>
__global__ void __launch_bounds__(BLOCK) k_refresh_lps(const FastArgs fa) {
  if (fa.gate && !fa.gate[(size_t)blockIdx.y * fa.gate_stride]) return;
  __shared__ FastH H;
  __shared__ MathTabs T;
  const SweepArgs &a = fa.s;
  const int chain = blockIdx.y;
  load_fast(H, T, fa.fasth + chain, fa.exp_tab, fa.log_tab);
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= a.G) return;
  const int64_t i = (int64_t)chain * a.G + g;
  const uint32_t f = a.flags[i];
  if (f & F_INSPIKE) { fa.lps[i] = 0; if (fa.want_stats && (f & F_ALLZERO)) a.xglik[i] = 0.0; return; }
  const double A = detect_logsum<RT>(H, T, fa.nd_mask[g], a.gl[g] * fexp(a.Yg[i], T.exp_t));
  fa.lps[i] = A;
  // want_stats doubles as "also rebuild XgLikelihood of the all-zero genes" (n_det = 0 => XgLik = beta A): zz_sweep_host does not
  // upload the caches, and the spike move reads this one (Pr:581)
  if (fa.want_stats && (f & F_ALLZERO)) a.xglik[i] = times_beta(H, A);
}

}  // namespace
#include "zz_persist.cuh"
namespace {

// before a segment of generations: proposes the first generation's mhP_x (Pr:193-243, first half) and derives the sweep constants
__global__ void k_gen_prepare(int C, int R, int tune, int hyper_moves, uint64_t seed, zz_hyper *hyper, ChainDev *chain, FastH *fasth) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  derive_fast(fasth[c], hyper + c);
  if (!hyper_moves) return;
  ChainDev &cd = chain[c];
  cd.px_prev_acc = 0;
  px_propose(cd, hyper[c], seed, cd.step, tune, R);
  set_px(fasth[c], hyper[c], cd);
}
// after a segment whose last generation accepted its mhP_x: A(y) += the stored per-gene term (inside a segment the next sweep does it)
__global__ void __launch_bounds__(BLOCK) k_apply_dlt(int64_t G, const ChainDev *chain, const uint16_t *flags, const double *dlt, double *lps) {
  if (!chain[blockIdx.y].px_prev_acc) return;
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= G) return;
  const int64_t i = (int64_t)blockIdx.y * G + g;
  if (!(flags[i] & F_INSPIKE)) lps[i] += dlt[i];
}

// ---- zz_sweep_host: the state of a CHUNK of genes (caller order [c0, c0 + n)) enters and leaves through staging planes, so
// that chunk k+1 can be on the PCIe bus (host -> device) and chunk k-1 on its way back while chunk k is being swept.

// inbound: scatter the chunk's state into its slots, pack the allocation words and rebuild what the production sweep carries
// but the host does not hold: A(Yg) (`lps`) and the XgLikelihood of all-zero genes (k_refresh_lps, chunk-wise)
template <int RT>
__global__ void __launch_bounds__(BLOCK) k_host_in(const FastArgs fa, const HostChunk hc) {
  __shared__ FastH H;
  __shared__ MathTabs T;
  const SweepArgs &a = fa.s;
  load_fast(H, T, fa.fasth, fa.exp_tab, fa.log_tab);
  const int64_t j = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (j >= hc.n) return;
  const int64_t c = hc.c0 + j, i = hc.inv[c];
  ZZ_DEV_ASSERT(i >= 0 && i < a.G);
  const double y = hc.sYg[c];
  a.Yg[i] = y; a.vg[i] = hc.svg[c]; hc.ty[i] = hc.sty[c]; hc.tv[i] = hc.stv[c];
  uint32_t f = a.flags[i];
  f = (f & ~F_ACTIVE) | (hc.sai[c] ? F_ACTIVE : 0u);
  f = flag_set_k(f, hc.sw[c] < 1 ? 1 : hc.sw[c]);
  f = (f & ~F_INSPIKE) | (hc.ss[c] ? F_INSPIKE : 0u);
  a.flags[i] = (uint16_t)f;
  if (f & F_INSPIKE) { fa.lps[i] = 0; if (f & F_ALLZERO) a.xglik[i] = 0.0; return; }
  const double A = detect_logsum<RT>(H, T, fa.nd_mask[i], a.gl[i] * fexp(y, T.exp_t));
  fa.lps[i] = A;
  if (f & F_ALLZERO) a.xglik[i] = times_beta(H, A);
}

// outbound: the chunk's updated state, caches and allocation vectors back into caller order
__global__ void __launch_bounds__(BLOCK) k_host_out(const SweepArgs a, const HostChunk hc) {
  const int64_t j = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (j >= hc.n) return;
  const int64_t c = hc.c0 + j, i = hc.inv[c];
  ZZ_DEV_ASSERT(i >= 0 && i < a.G);
  hc.dYg[c] = a.Yg[i]; hc.dvg[c] = a.vg[i]; hc.dxl[c] = a.xglik[i]; hc.dvp[c] = a.vgprob[i];
  const uint32_t f = a.flags[i];
  hc.dai[c] = (f & F_ACTIVE) ? 1 : 0; hc.dw[c] = flag_k(f); hc.ds[c] = (f & F_INSPIKE) ? 1 : 0;
}

// exports the Philox draws of one move slot in the layout of that move's `draws` argument
__global__ void __launch_bounds__(BLOCK) k_draws(int64_t G, int64_t gene_offset, const int32_t *perm, int C, uint64_t seed, uint64_t step, int slot, double *out) {
  const int chain = blockIdx.y;
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= G) return;
  const int64_t i = (int64_t)chain * G + g, CG = (int64_t)C * G;
  const uint64_t gid = (uint64_t)(gene_offset + perm[g]);
  uint32_t b0[4], b1[4], b2[4];
  double z0, z1;
  switch (slot) {
    case ZZ_SLOT_YG:
      philox_block(seed, chain, step, 0, gid, b0); box_muller(u32_open(b0[0]), u32_open(b0[1]), z0, z1);
      out[i] = z0; out[CG + i] = u32_open(b0[2]); break;
    case ZZ_SLOT_VARIANCE_G:
      philox_block(seed, chain, step, 0, gid, b0); philox_block(seed, chain, step, 1, gid, b1);
      out[i] = u32_open(b0[3]); out[CG + i] = u32_open(b1[0]); break;
    case ZZ_SLOT_ALLOC:
      philox_block(seed, chain, step, 1, gid, b1);
      out[i] = u32_open(b1[1]); out[CG + i] = u32_open(b1[2]); break;
    case ZZ_SLOT_ALLOC_WITHIN:
      philox_block(seed, chain, step, 1, gid, b1);
      out[i] = u32_open(b1[2]); break;
    case ZZ_SLOT_SPIKE:
      philox_block(seed, chain, step, 2, gid, b2); box_muller(u32_open(b2[0]), u32_open(b2[1]), z0, z1);
      out[i] = z0; out[CG + i] = z1; out[2 * CG + i] = u32_open(b2[2]); break;
  }
}

enum { EVAL_XGLIK = 0, EVAL_VGPRIOR = 1, EVAL_LEVEL2 = 2, EVAL_REFRESH = 3, EVAL_COMMIT_VG = 4 };

struct EvalArgs {
  int64_t G; int which;
  const double *Xg, *gl, *Yg, *vg;
  const uint16_t *flags;
  const zz_hyper *hyper;   // already offset to the chain (or base when grid.y spans chains)
  double *out0, *out1;     // per-gene outputs (chain-offset applied by blockIdx.y * G)
  const int *gate;         // device-resident loop: chain c does nothing unless gate[c * gate_stride] != 0 (an accepted move)
  int gate_stride;
};

__global__ void __launch_bounds__(BLOCK) k_eval(const EvalArgs a) {
  if (a.gate && !a.gate[(size_t)blockIdx.y * a.gate_stride]) return;
  __shared__ HyperS H;
  const int chain = blockIdx.y;
  load_hyper(H, a.hyper + chain);
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= a.G) return;
  const int64_t i = (int64_t)chain * a.G + g;
  const double y = a.Yg[i], v = a.vg[i];
  const uint32_t f = a.flags[i];
  const bool in_spike = f & F_INSPIKE;
  if (a.which == EVAL_XGLIK || a.which == EVAL_REFRESH) {
    a.out0[i] = in_spike ? xg_lik_inspike(H, f & F_ALLZERO) : xg_lik_out(H, a.Xg + g, a.G, y, v, a.gl[g]);
  }
  if (a.which == EVAL_VGPRIOR) a.out0[i] = varg_prior(H, v, get_sigmax(H.s0, H.s1, y), H.tau, H.rtau);
  if (a.which == EVAL_REFRESH) a.out1[i] = varg_prior(H, v, get_sigmax(H.s0, H.s1, y), H.tau, H.rtau);   // I:504
  if (a.which == EVAL_COMMIT_VG && !in_spike) a.out1[i] = varg_prior(H, v, get_sigmax(H.s0, H.s1, y), H.tau, H.rtau);  // Pr:113,147,187
  if (a.which == EVAL_LEVEL2) a.out0[i] = level2_lik(H, y, f, in_spike);
}

enum { RED_VARG_HYPER = 0, RED_LEVEL2 = 1, RED_PXDELTA = 2, RED_LNL = 3 };

struct RedArgs {
  int64_t G; int op, chain;
  const double *Xg, *gl, *Yg, *vg, *xglik, *vgprob;  // chain-offset applied by the launcher
  const uint16_t *flags;
  const zz_hyper *hyper;                              // this chain's
  double s0, s1, tau, im, iv;
  double am[ZZ_MAX_COMP], av[ZZ_MAX_COMP];
  double alpha_prop[ZZ_MAX_LIBS];
  uint8_t lib_mask[ZZ_MAX_LIBS];
  double *partials;                                   // [grid.y][grid.x][2]
};

// Grid-stride per-gene reduction; blockIdx.y = library for RED_PXDELTA.  Two sums per gene at most.
__global__ void __launch_bounds__(BLOCK) k_reduce(const RedArgs a) {
  __shared__ HyperS H;
  __shared__ double red[BLOCK / 32];
  // blockIdx.z = chain of a batched launch (the sample rows of zz_run_generations); the host entry points launch with grid.z = 1 and
  // pointers already offset to their chain
  const int64_t coff = (int64_t)blockIdx.z * a.G;
  const double *Yg = a.Yg + coff, *vg = a.vg + coff, *vgprob = a.vgprob + coff;
  const uint16_t *flags = a.flags + coff;
  load_hyper(H, a.hyper + blockIdx.z);
  const int lib = (int)blockIdx.y;
  double v0 = 0, v1 = 0;
  const bool active_y = !(a.op == RED_PXDELTA && !a.lib_mask[lib]);
  const double p_s0 = a.s0, p_s1 = a.s1, p_tau = a.tau, p_im = a.im, p_iv = a.iv;
  const double *p_am = a.am, *p_av = a.av;
  const double p_alpha = a.op == RED_PXDELTA ? a.alpha_prop[lib] : 0.0;
  if (active_y) {
    for (int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.G; g += (int64_t)gridDim.x * BLOCK) {
      const double y = Yg[g], v = vg[g];
      const uint32_t f = flags[g];
      const bool in_spike = f & F_INSPIKE;
      if (a.op == RED_VARG_HYPER) {                  // Pr:84-87, 126-130, 167-170
        if (!in_spike) {
          v0 += varg_prior(H, v, get_sigmax(p_s0, p_s1, y), p_tau, sqrt(p_tau));
          v1 += vgprob[g];
        }
      } else if (a.op == RED_LEVEL2) {               // Pr:444-447, 653-654, 780-784
        if (f & F_ACTIVE) {
          const int k = flag_k(f) - 1;
          const double d = y - p_am[k];
          v1 += -log(H.sqrt2pi * sqrt(p_av[k])) - (d * d) / (2 * p_av[k]);
        } else {
          v0 += inactive_lik_at(H, y, in_spike, p_im, p_iv);
        }
      } else if (a.op == RED_PXDELTA) {              // Pr:209-224
        const double x = a.Xg[(int64_t)lib * a.G + g];
        v0 += xg_lik_onelib(H, x, y, v, a.gl[g], p_alpha, in_spike, f & F_ALLZERO) -
              xg_lik_onelib(H, x, y, v, a.gl[g], H.alpha[lib], in_spike, f & F_ALLZERO);
      } else {                                       // RED_LNL, U:70-75
        v0 += in_spike ? xg_lik_inspike(H, f & F_ALLZERO) : xg_lik_out(H, a.Xg + g, a.G, y, v, a.gl[g]);
      }
    }
  }
  const double s0 = block_sum<BLOCK>(v0, red);
  const double s1 = block_sum<BLOCK>(v1, red);
  if (threadIdx.x == 0) {
    double *p = a.partials + (((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 2;
    p[0] = s0; p[1] = s1;
  }
}

// partials [ny][nblk][nval] -> out [ny][nval], fixed summation order: warp w owns values w, w + 8, ...; lane l adds rows
// l, l + 32, ... then a butterfly — no block barrier, so the whole reduction is one short latency chain
__global__ void __launch_bounds__(BLOCK) k_reduce_final(const double *partials, int nblk, int nval, double *out) {
  __shared__ double vec[MAX_NSTAT + RED_THREADS];
  const int y = blockIdx.x;                          // one chain (or reduction slot) per block, canonical order of reduce_rows_to_smem
  reduce_rows_to_smem(partials + (int64_t)y * nblk * nval, nblk, nval, nval, vec);
  if ((int)threadIdx.x < nval) out[(int64_t)y * nval + threadIdx.x] = vec[threadIdx.x];
}




struct StatArgs {
  int64_t G; int K;
  const double *Yg, *vg, *xglik, *vgprob;  // base pointers, chain = blockIdx.y
  const uint16_t *flags;
  double *partials;                         // [C][nblk][nstat]
};

// sufficient statistics for the hyper moves (layout documented at zz_suffstats in the header)
__global__ void __launch_bounds__(BLOCK) k_suffstats(const StatArgs a) {
  __shared__ double red[BLOCK / 32];
  const int chain = blockIdx.y, K1 = a.K + 1, nstat = 3 * K1 + NSTAT_FIXED;
  double cn[ZZ_MAX_COMP + 1], cy[ZZ_MAX_COMP + 1], cy2[ZZ_MAX_COMP + 1], fx[NSTAT_FIXED];
#pragma unroll
  for (int c = 0; c <= ZZ_MAX_COMP; ++c) { cn[c] = 0; cy[c] = 0; cy2[c] = 0; }
#pragma unroll
  for (int j = 0; j < NSTAT_FIXED; ++j) fx[j] = 0;
  for (int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.G; g += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = (int64_t)chain * a.G + g;
    const double y = a.Yg[i], v = a.vg[i];
    const uint32_t f = a.flags[i];
    const bool act = f & F_ACTIVE, sp = f & F_INSPIKE;
    const int comp = act ? flag_k(f) : 0;                          // Pr:415 all_allocations
    const bool use_y = act || !sp;
#pragma unroll
    for (int c = 0; c <= ZZ_MAX_COMP; ++c) {
      if (c == comp) { cn[c] += 1; if (use_y) { cy[c] += y; cy2[c] += y * y; } }
    }
    fx[0] += sp ? 1.0 : 0.0;
    if (!sp) {
      const double z = log(v);
      fx[1] += 1; fx[2] += z; fx[3] += z * z; fx[4] += y; fx[5] += y * y; fx[6] += y * z;
      fx[8] += a.vgprob[i];
    }
    fx[7] += a.xglik[i];
    if (!isfinite(y) || !isfinite(v)) fx[9] += 1;
  }
  double *p = a.partials + ((int64_t)chain * gridDim.x + blockIdx.x) * nstat;
#pragma unroll
  for (int c = 0; c <= ZZ_MAX_COMP; ++c) {
    if (c < K1) {
      const double t0 = block_sum<BLOCK>(cn[c], red), t1 = block_sum<BLOCK>(cy[c], red), t2 = block_sum<BLOCK>(cy2[c], red);
      if (threadIdx.x == 0) { p[c] = t0; p[K1 + c] = t1; p[2 * K1 + c] = t2; }
    }
  }
#pragma unroll
  for (int j = 0; j < NSTAT_FIXED; ++j) {
    const double t = block_sum<BLOCK>(fx[j], red);
    if (threadIdx.x == 0) p[3 * K1 + j] = t;
  }
}

// accepted mhP_x with R > 2: XgLikelihood += ll_lib(alpha_new) - ll_lib(alpha_old), Pr:213-215, 239
__global__ void __launch_bounds__(BLOCK) k_px_commit(int64_t G, int lib, double alpha_new, const double *Xg, const double *gl,
                                                      const double *Yg, const double *vg, const uint16_t *flags,
                                                      const zz_hyper *hyper, double *xglik) {
  __shared__ HyperS H;
  load_hyper(H, hyper);
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= G) return;
  const uint32_t f = flags[g];
  const double x = Xg[(int64_t)lib * G + g];
  xglik[g] = xglik[g] + xg_lik_onelib(H, x, Yg[g], vg[g], gl[g], alpha_new, f & F_INSPIKE, f & F_ALLZERO) -
             xg_lik_onelib(H, x, Yg[g], vg[g], gl[g], H.alpha[lib], f & F_INSPIKE, f & F_ALLZERO);
}

// one sample: the _model_parameters.log row (thread 0) and, when due, the candidate genes' Yg / variance_g rows, written straight
// into pinned host memory (zz_set_sample_sink)
__global__ void __launch_bounds__(BLOCK) k_sample_row(double gen, const zz_hyper *hy, const double *lnl, int R, int K, double *prow,
                                                       const int32_t *cand_slots, int64_t n_cand, const double *Yg, const double *vg,
                                                       const uint16_t *flags, double *ygrow, double *vgrow) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int j = 0;
    prow[j++] = gen; prow[j++] = lnl[0]; prow[j++] = hy->s0; prow[j++] = hy->s1; prow[j++] = hy->tau;
    for (int r = 0; r < R; ++r) prow[j++] = hy->alpha_r[r];
    prow[j++] = hy->spike_probability; prow[j++] = hy->inactive_means; prow[j++] = hy->inactive_variances;
    for (int k = 0; k < K; ++k) prow[j++] = hy->active_means[k];
    for (int k = 0; k < K; ++k) prow[j++] = hy->active_variances[k];
    prow[j++] = hy->weight_active;
    for (int k = 0; k < K; ++k) prow[j++] = hy->weight_within_active[k];
    if (ygrow) { ygrow[0] = gen; vgrow[0] = gen; }
  }
  if (!ygrow) return;
  for (int64_t c = (int64_t)blockIdx.x * BLOCK + threadIdx.x; c < n_cand; c += (int64_t)gridDim.x * BLOCK) {
    const int64_t i = cand_slots[c];
    ZZ_DEV_ASSERT(i >= 0);
    const bool sp = flags[i] & F_INSPIKE;
    ygrow[1 + c] = sp ? -10.0 : Yg[i];                  // FileWriting:93: Yg * (1 - spike) + (-10) * spike
    vgrow[1 + c] = sp ? 0.0 : vg[i];                    // FileWriting:100
  }
}

// ---- posterior-predictive discrepancies (R/zigzagPostPredictiveSimulation.R:3-156) -------------------------
// One simulated data set per call: x_sim[g, r] ~ N(Yg, variance_g), dropped with probability 1 - p_x[g, r], all -Inf for in-spike
// genes; Yg_sim from the gene's mixture component.  The reference's statistics are functionals of binned cumulative
// distributions, so the O(G R) work is histograms: integer counts through atomics (exact, order-free), the p_x-weighted model
// histogram in 2^-40 fixed point (exact sums again; quantisation <= 4.5e-13 relative at 1 M genes), and three per-library
// floating-point sums through fixed-order block partials.  Pass 2 repeats the (counter-based) draws to bin the mean-centred copies.
struct PPArgs {
  int64_t G; int R, nb, pass;
  uint64_t seed, step; int64_t gene_offset;
  const double *Xg, *gl, *Yg, *vg; const uint16_t *flags; const int32_t *perm; const zz_hyper *hyper; const double *bins;
  unsigned *cd, *cs, *cds, *css, *hy, *hsy;          // [R][nb+1] x 4, [nb+1] x 2
  unsigned long long *md;                            // [R][nb+1], 2^-40 fixed point
  double *partials;                                  // [R][gridDim.x][4]: sum x (data), sum x (sim), sum (1 - p_x) over out-of-spike genes
  double shift_d[ZZ_MAX_LIBS], shift_s[ZZ_MAX_LIBS]; // pass 2: grand mean - library mean
};
__device__ __forceinline__ int upper_idx(const double *bins, int nb, double x) {   // number of bins <= x: (x < bins[b]) <=> b >= idx
  int lo = 0, hi = nb;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (bins[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
__global__ void __launch_bounds__(BLOCK) k_post_pred(const PPArgs a) {
  // one library per blockIdx.y: the CTA's histograms of that library live in shared memory and reach global memory as one atomic
  // per non-empty bin at the end (48 M global atomics per call at 1 M x 16 otherwise)
  extern __shared__ double sbins[];                   // [nb] bins, then [nb+1] u64 model weights, then 4 x [nb+1] u32 counts
  __shared__ HyperS H;
  __shared__ double red[BLOCK / 32];
  const int r = blockIdx.y, Hn = a.nb + 1;
  unsigned long long *s_md = reinterpret_cast<unsigned long long *>(sbins + a.nb);
  unsigned *s_c0 = reinterpret_cast<unsigned *>(s_md + Hn), *s_c1 = s_c0 + Hn, *s_hy = s_c1 + Hn, *s_hsy = s_hy + Hn;
  load_hyper(H, a.hyper);
  for (int b = threadIdx.x; b < a.nb; b += BLOCK) sbins[b] = a.bins[b];
  for (int b = threadIdx.x; b < Hn; b += BLOCK) { s_md[b] = 0; s_c0[b] = 0; s_c1[b] = 0; s_hy[b] = 0; s_hsy[b] = 0; }
  __syncthreads();
  double sxd = 0, sxs = 0, s1mp = 0;
  for (int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.G; g += (int64_t)gridDim.x * BLOCK) {
    const double y = a.Yg[g], sd = sqrt(a.vg[g]);
    const uint32_t f = a.flags[g];
    const bool sp = f & F_INSPIKE;
    const uint64_t gid = (uint64_t)(a.gene_offset + a.perm[g]);
    const double px = 1 - exp(-((a.gl[g] * exp(y)) * H.alpha[r]));                    // U:150-159
    const double x = a.Xg[(int64_t)r * a.G + g];
    uint32_t w[4];
    philox_block(a.seed, 0, a.step, 16 + r, gid, w);
    double z0, z1;
    box_muller(u32_open(w[0]), u32_open(w[1]), z0, z1);
    double sx = y + sd * z0;                                                          // PostPred:24
    if (u32_open(w[2]) < 1 - px) sx = -INFINITY;                                      // PostPred:30-31
    if (sp) sx = -INFINITY;                                                           // PostPred:38
    if (a.pass == 1) {
      if (x != -INFINITY) { atomicAdd(s_c0 + upper_idx(sbins, a.nb, x), 1u); sxd += x; }
      if (sx != -INFINITY) { atomicAdd(s_c1 + upper_idx(sbins, a.nb, sx), 1u); sxs += sx; }
      if (!sp) {
        const int iy = upper_idx(sbins, a.nb, y);
        atomicAdd(s_md + iy, (unsigned long long)(px * 1099511627776.0 + 0.5));
        s1mp += 1 - px;
        if (r == 0) {
          atomicAdd(s_hy + iy, 1u);
          philox_block(a.seed, 0, a.step, 15, gid, w);
          box_muller(u32_open(w[0]), u32_open(w[1]), z0, z1);
          const double sy = (f & F_ACTIVE) ? H.am[flag_k(f) - 1] + sqrt(0.5 * H.two_av[flag_k(f) - 1]) * z0 : H.im + H.sd_i * z0;   // PostPred:79-84
          atomicAdd(s_hsy + upper_idx(sbins, a.nb, sy), 1u);
        }
      }
    } else {
      if (x != -INFINITY) atomicAdd(s_c0 + upper_idx(sbins, a.nb, x + a.shift_d[r]), 1u);      // PostPred:52-58
      if (sx != -INFINITY) atomicAdd(s_c1 + upper_idx(sbins, a.nb, sx + a.shift_s[r]), 1u);
    }
  }
  __syncthreads();
  unsigned *g0 = (a.pass == 1 ? a.cd : a.cds) + (size_t)r * Hn, *g1 = (a.pass == 1 ? a.cs : a.css) + (size_t)r * Hn;
  for (int b = threadIdx.x; b < Hn; b += BLOCK) {
    if (s_c0[b]) atomicAdd(g0 + b, s_c0[b]);
    if (s_c1[b]) atomicAdd(g1 + b, s_c1[b]);
    if (a.pass == 1) {
      if (s_md[b]) atomicAdd(a.md + (size_t)r * Hn + b, s_md[b]);
      if (r == 0) { if (s_hy[b]) atomicAdd(a.hy + b, s_hy[b]); if (s_hsy[b]) atomicAdd(a.hsy + b, s_hsy[b]); }
    }
  }
  if (a.pass == 1) {
    const double t0 = block_sum<BLOCK>(sxd, red), t1 = block_sum<BLOCK>(sxs, red), t2 = block_sum<BLOCK>(s1mp, red);
    if (threadIdx.x == 0) { double *p = a.partials + ((int64_t)r * gridDim.x + blockIdx.x) * 4; p[0] = t0; p[1] = t1; p[2] = t2; p[3] = 0; }
  }
}

// allocation_trace += one-hot(allocation), M:194-199
__global__ void __launch_bounds__(BLOCK) k_alloc_trace(int64_t G, int K, const uint16_t *flags, uint32_t *trace) {
  const int chain = blockIdx.y;
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= G) return;
  const uint32_t f = flags[(int64_t)chain * G + g];
  const int comp = (f & F_ACTIVE) ? flag_k(f) : 0;
  ZZ_DEV_ASSERT(comp >= 0 && comp <= K);
  trace[((int64_t)chain * (K + 1) + comp) * G + g] += 1u;
}

// Constructor-side initialisation of the per-gene state, R/zigzagInitialize.R:416-504, for every chain of the handle from that
// chain's hyper-parameters.  Draw map (Philox, step = the caller's): block 0 = {u of the active/inactive allocation (I:367), u of
// the within-active allocation (I:372), Box-Muller pair (z for rwm of an all-zero gene I:431-438, z for rwv of a gene with < 2
// detected libraries I:441-448)}; block 1 = pair (z of Yg I:449, z of variance_g I:465); block 2 + j = pair of re-draw j of a gene
// whose likelihood is 0 (I:480-502; each gene repeats on its own until its likelihood is positive, at most 1000 times — the
// reference's loop re-draws exactly the genes that are still at -Inf, so the per-gene loop visits the same states).
// rwm / rwv come from the per-gene data statistics (xbar, xss, nd_mask); `rv_m`, `rv_s` = log(mean) and sd of the row variances of
// the fully detected genes of the WHOLE data set (host, I:446).
__global__ void __launch_bounds__(BLOCK) k_init_state(int64_t G, int64_t gene_offset, uint64_t seed, uint64_t step, double rv_m, double rv_s,
                                                       const double *__restrict__ Xg, const double *__restrict__ gl, const int32_t *__restrict__ perm,
                                                       const uint64_t *__restrict__ nd_mask, const double *__restrict__ xbar, const double *__restrict__ xss,
                                                       const zz_hyper *hyper, double *Yg, double *vg, double *ty, double *tv, uint16_t *flags) {
  __shared__ HyperS H;
  const int chain = blockIdx.y;
  load_hyper(H, hyper + chain);
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (g >= G) return;
  const int64_t i = (int64_t)chain * G + g;
  const uint64_t gid = (uint64_t)(gene_offset + perm[g]);
  const int n_det = H.R - __popcll(nd_mask[g]);
  const bool allzero = n_det == 0;
  uint32_t f = flags[i] & F_PINNED;
  if (allzero) f |= F_ALLZERO | F_INSPIKE;                            // I:415-418
  uint32_t b[4];
  philox_block(seed, chain, step, 0, gid, b);
  const double u_ai = u32_open(b[0]), u_w = u32_open(b[1]);
  double z_m, z_v0, z_y, z_v;
  box_muller(u32_open(b[2]), u32_open(b[3]), z_m, z_v0);
  if (((f & F_PINNED) || u_ai < H.wa) && !allzero) f |= F_ACTIVE;     // I:367-369, I:422
  int k = 1; double cum = 0;
  for (int j = 0; j < H.K - 1; ++j) { cum += H.w[j]; k += (cum <= u_w); }   // I:372
  f = flag_set_k(f, k);
  philox_block(seed, chain, step, 1, gid, b);
  box_muller(u32_open(b[0]), u32_open(b[1]), z_y, z_v);
  const double rwm = n_det > 0 ? xbar[g] : H.im + H.sd_i * z_m;       // I:431-438
  const double rwv = n_det >= 2 ? xss[g] / (n_det - 1) : exp(rv_m + rv_s * z_v0);   // I:441-448
  double y = rwm + sqrt(rwv) * z_y;                                   // I:449
  double v = exp(0.5 + 0.2 * z_v);                                    // I:465
  if (v > H.vub) v = H.vub;                                           // I:468
  if (!allzero) {                                                     // I:480-502
    for (uint32_t j = 0; j < 1000 && xg_lik_out(H, Xg + g, G, y, v, gl[g]) == -INFINITY; ++j) {
      double z1, z2;
      philox_block(seed, chain, step, 2 + j, gid, b);
      box_muller(u32_open(b[0]), u32_open(b[1]), z1, z2);
      y = H.im + H.sd_i * z1;
      v = exp(log(exp(H.s0 + H.s1 * y)) + H.tau + H.rtau * z2);
      if (v > H.vub) v = H.vub;
    }
  }
  Yg[i] = y; vg[i] = v; ty[i] = 0.5; tv[i] = 0.5;                      // tuning parameters I:185-199
  flags[i] = (uint16_t)f;
}

// x_tune on the per-gene windows, U:161-181 with the clamps of U:202-206
__device__ __forceinline__ double x_tune(double wsum, double t, double target, double lo, double hi) {
  const double pjump = wsum / 100;
  double nt = t * (tan(PI_D * pjump / 2) / tan(PI_D * target / 2));
  if (nt < lo) nt = lo;
  if (nt > hi) nt = hi;
  return nt;
}
__global__ void __launch_bounds__(BLOCK) k_tune(int64_t n, double target, const ulonglong2 *win_y, const ulonglong2 *win_v, double *ty, double *tv) {
  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  tv[i] = x_tune(window_sum(win_v[i]), tv[i], target, 0.0001, 10);
  ty[i] = x_tune(window_sum(win_y[i]), ty[i], target, 0.0001, 10);
}
__global__ void __launch_bounds__(BLOCK) k_window_init(int64_t n, ulonglong2 *win_y, ulonglong2 *win_v) {
  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const ulonglong2 w = make_ulonglong2((1ULL << 23) - 1, 0ULL);   // c(rep(0,77), rep(1,23)), I:460,466
  win_y[i] = w; win_v[i] = w;
}
__global__ void __launch_bounds__(BLOCK) k_window_sums(int64_t n, const ulonglong2 *win_y, const ulonglong2 *win_v, int32_t *sy, int32_t *sv) {
  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  sy[i] = window_sum(win_y[i]); sv[i] = window_sum(win_v[i]);
}


// dst[p][j] = src[p][perm[j]]  (caller order -> device order) and its inverse; grid.y = plane
template <class T>
__global__ void __launch_bounds__(BLOCK) k_gather(int64_t n, const int32_t *__restrict__ perm, const T *__restrict__ src, T *__restrict__ dst) {
  const int64_t j = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (j < n) { ZZ_DEV_ASSERT(perm[j] >= 0 && perm[j] < n); dst[(int64_t)blockIdx.y * n + j] = src[(int64_t)blockIdx.y * n + perm[j]]; }
}
template <class T>
__global__ void __launch_bounds__(BLOCK) k_scatter(int64_t n, const int32_t *__restrict__ perm, const T *__restrict__ src, T *__restrict__ dst) {
  const int64_t j = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (j < n) { ZZ_DEV_ASSERT(perm[j] >= 0 && perm[j] < n); dst[(int64_t)blockIdx.y * n + perm[j]] = src[(int64_t)blockIdx.y * n + j]; }
}

// int32 R-style allocation vectors (caller order) <-> packed flag word (device order)
__global__ void __launch_bounds__(BLOCK) k_pack_flags(int64_t n, const int32_t *perm, const int32_t *ai, const int32_t *within, const int32_t *spike,
                                                       const int32_t *pinned, uint16_t *flags) {
  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const int32_t o = perm[i];
  ZZ_DEV_ASSERT(o >= 0 && o < n);
  uint32_t f = flags[i];
  if (ai) f = (f & ~F_ACTIVE) | (ai[o] ? F_ACTIVE : 0u);
  if (within) f = flag_set_k(f, within[o] < 1 ? 1 : within[o]);
  if (spike) f = (f & ~F_INSPIKE) | (spike[o] ? F_INSPIKE : 0u);
  if (pinned) f = (f & ~F_PINNED) | (pinned[o] ? F_PINNED : 0u);
  flags[i] = (uint16_t)f;
}
__global__ void __launch_bounds__(BLOCK) k_unpack_flags(int64_t n, const int32_t *perm, const uint16_t *flags, int32_t *ai, int32_t *within, int32_t *spike) {
  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const int32_t o = perm[i];
  ZZ_DEV_ASSERT(o >= 0 && o < n);
  const uint32_t f = flags[i];
  if (ai) ai[o] = (f & F_ACTIVE) ? 1 : 0;
  if (within) within[o] = flag_k(f);
  if (spike) spike[o] = (f & F_INSPIKE) ? 1 : 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ handle

#define ZZ_HOST_MAX_CHUNKS 16

struct zz_handle {
  zz_config cfg;
  int64_t G; int C, R, K;
  cudaStream_t stream; bool own_stream;
  std::string err; bool sticky;
  uint64_t launches;
  // device
  double *Xg, *gl, *Yg, *vg, *xglik, *vgprob, *ty, *tv;
  uint16_t *flags;
  ulonglong2 *win_y, *win_v;
  uint32_t *trace;
  zz_hyper *hyper;
  double *partials, *red_out, *scratch;   // scratch: 9*C*G doubles for draw tables / eval output
  uint8_t *dec;                            // 4*C*G
  int32_t *iscratch;                       // 3*C*G
  int *nan_flag;
  bool have_data;
  double *lps; double *exp_tab; double2 *log_tab; int64_t n_allzero;
  int32_t *perm; std::vector<int32_t> perm_h;   // device slot -> caller's gene index (expression-sorted, all-zero genes last)
  double *stage;                           // staging for permuted transfers, same size as scratch
  uint64_t *nd_mask; double *xbar, *xss;   // per-gene data statistics (k_prepare_data)
  FastH *fasth; bool fasth_valid;           // derived hyper-parameters of the production kernels
  // peer-memory all-reduce of the statistics vector (zz_comm_*)
  int comm_rank, comm_world; unsigned long long comm_epoch, comm_epoch_small; bool comm_ready; long long comm_timeout_ns;
  double *comm_inbox; unsigned long long *comm_flags;        // this rank's inbox / flags (device memory, exported through CUDA IPC)
  double *peer_inbox[8]; unsigned long long *peer_flags[8];  // every rank's buffers as mapped here (own entry = local pointers)
  unsigned *tile_counter; int persist_blocks;       // tile_counter: ticket of the "last CTA done" epilogue
  double *stats_target; int stats_allreduce;   // zz_set_stats_output_dev: where the sweep's fused epilogue leaves the statistics
  double *sw_partials; int64_t sw_rows; bool stats_fresh;   // statistics emitted by the last production sweep
  bool lps_valid;                          // lps == A(current Yg, alpha_r) for every out-of-spike gene
  std::vector<zz_hyper> hyper_host;
  double *pinned;                          // small pinned staging for reduction results
  char *bounce; size_t bounce_cap;         // pinned bounce region for state downloads into pageable caller memory (zz_get_state)
  cudaEvent_t bounce_ev[9]; bool bounce_ev_ok;
  // zz_sweep_host pipeline (created on first use)
  void *pp_buf; int pp_nb;                 // zz_post_predictive: histograms + bins on the device
  bool sink_on; zz_sample_sink sink; double *sink_p, *sink_y, *sink_v; int32_t *cand_slots; int64_t sink_prow, sink_yrow;   // zz_set_sample_sink
  bool hyper_host_stale;                   // zz_run_generations moved the hyper-parameters on the device; hyper_host is refreshed by zz_chain_read
  ChainDev *devchain; double *chain_stats; bool chain_ready; int64_t chain_gen, chain_samples; uint64_t chain_step;   // zz_run_generations
  double *dlt; GenCtl *genctl; long long *genacc; int gen_blocks, gen_blocks_l, gen_blocks_c;   // k_generations: per-gene mhP_x terms, grid-barrier words, fixed-point totals
  bool caches_valid;                       // XgLikelihood / variance_g_probability are current (device-resident segments do not maintain them)
  int32_t *inv_perm; cudaStream_t s_up, s_down; cudaEvent_t ev_entry, ev_up[ZZ_HOST_MAX_CHUNKS], ev_done[ZZ_HOST_MAX_CHUNKS]; bool pipe_ready;
};

namespace {

int fail(zz_handle *h, int code, const std::string &msg) {
  if (h) { h->err = msg; if (code == ZZ_ERR_CUDA) h->sticky = true; }
  return code;
}

// Takes `n` consecutive exchange epochs; returns the one BEFORE the first.  The low 32 bits travel with every word and mark it
// valid, so they must never be 0 (the zero-initialised inbox): a range that would cross a multiple of 2^32 is moved behind it,
// keeping the slot parity alternating (slot = epoch & 1; two exchanges in a row on one slot could overwrite unread words).
// Every rank calls this with the same sequence of `n`, so the epochs stay in step.
unsigned long long comm_take_epochs(zz_handle *h, unsigned long long n) {
  if ((h->comm_epoch & 0xFFFFFFFFull) + n > 0xFFFFFFFFull) {
    const unsigned long long last = h->comm_epoch;
    h->comm_epoch = (h->comm_epoch | 0xFFFFFFFFull) + 1ull;          // low half 0: the first epoch used is this + 1
    if (((h->comm_epoch + 1ull) & 1ull) == (last & 1ull)) h->comm_epoch += 1ull;
  }
  const unsigned long long before = h->comm_epoch;
  h->comm_epoch += n;
  return before;
}
void fill_comm(zz_handle *h, CommArgs &cm, int nvec, unsigned long long epoch) {
  cm.rank = h->comm_rank; cm.world = h->comm_world; cm.nvec = nvec; cm.epoch = epoch; cm.base = 0; cm.stride = 2 * nvec;
  cm.timeout_ns = h->comm_timeout_ns;
  for (int r = 0; r < 8; ++r) { cm.inbox[r] = h->peer_inbox[r]; cm.flags[r] = h->peer_flags[r]; }
}
#define ZZ_CUDA(h, expr)                                                                         \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(h, ZZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
  } while (0)
#define ZZ_CHECK(h)                                                                              \
  do {                                                                                           \
    if (!(h)) return ZZ_ERR_INVALID;                                                             \
    if ((h)->sticky) return ZZ_ERR_CUDA;                                                         \
    ZZ_CUDA(h, cudaSetDevice((h)->cfg.device));                                                  \
  } while (0)
#define ZZ_LAUNCHED(h)                                                                           \
  do { (h)->launches++; ZZ_CUDA(h, cudaGetLastError()); } while (0)

inline dim3 gene_grid(const zz_handle *h, int64_t n, int ny = 1) { return dim3((unsigned)((n + BLOCK - 1) / BLOCK), (unsigned)ny); }
inline int red_blocks(int64_t n) { int64_t b = (n + BLOCK - 1) / BLOCK; return (int)(b < RED_BLOCKS ? (b < 1 ? 1 : b) : RED_BLOCKS); }

int ensure_caches(zz_handle *h);   // XgLikelihood / variance_g_probability current again after a device-resident segment

int chain_ok(zz_handle *h, int chain) {
  if (chain < 0 || chain >= h->C) return fail(h, ZZ_ERR_INVALID, "chain index out of range");
  return ZZ_OK;
}

template <class T> int up(zz_handle *h, T *dst, const T *src, int64_t n) {
  if (!src) return ZZ_OK;
  ZZ_CUDA(h, cudaMemcpyAsync(dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, h->stream));
  return ZZ_OK;
}
template <class T> int down(zz_handle *h, T *dst, const T *src, int64_t n) {
  if (!dst) return ZZ_OK;
  ZZ_CUDA(h, cudaMemcpyAsync(dst, src, sizeof(T) * n, cudaMemcpyDeviceToHost, h->stream));
  return ZZ_OK;
}

// per-gene arrays cross the ABI in the caller's gene order and live on the device in sorted order
template <class T> int up_g(zz_handle *h, T *dst, const T *src, int64_t n, int planes = 1) {
  if (!src) return ZZ_OK;
  T *st = reinterpret_cast<T *>(h->stage);
  ZZ_CUDA(h, cudaMemcpyAsync(st, src, sizeof(T) * n * planes, cudaMemcpyHostToDevice, h->stream));
  k_gather<T><<<dim3((unsigned)((n + BLOCK - 1) / BLOCK), (unsigned)planes), BLOCK, 0, h->stream>>>(n, h->perm, st, dst);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}
template <class T> int down_g(zz_handle *h, T *dst, const T *src, int64_t n, int planes = 1, bool sync = true) {
  if (!dst) return ZZ_OK;
  T *st = reinterpret_cast<T *>(h->stage);
  k_scatter<T><<<dim3((unsigned)((n + BLOCK - 1) / BLOCK), (unsigned)planes), BLOCK, 0, h->stream>>>(n, h->perm, src, st);
  ZZ_LAUNCHED(h);
  ZZ_CUDA(h, cudaMemcpyAsync(dst, st, sizeof(T) * n * planes, cudaMemcpyDeviceToHost, h->stream));
  if (sync) ZZ_CUDA(h, cudaStreamSynchronize(h->stream));   // the caller reads `dst` on return (stream order protects the staging buffer)
  return ZZ_OK;
}

void fill_sweep_args(zz_handle *h, SweepArgs &a, uint64_t step, int tune, uint32_t moves, const double *const draws_dev[9],
                     uint8_t *const dec_dev[4], int64_t g_begin, int64_t g_end) {
  a.G = h->G; a.g_begin = g_begin; a.g_end = g_end; a.gene_offset = h->cfg.gene_offset;
  a.C = h->C; a.R = h->R; a.K = h->K; a.tune = tune; a.moves = moves; a.seed = h->cfg.seed; a.step = step;
  a.Xg = h->Xg; a.gl = h->gl; a.Yg = h->Yg; a.vg = h->vg; a.xglik = h->xglik; a.vgprob = h->vgprob; a.ty = h->ty; a.tv = h->tv;
  a.flags = h->flags; a.win_y = h->win_y; a.win_v = h->win_v; a.hyper = h->hyper; a.perm = h->perm;
  for (int i = 0; i < 9; ++i) a.draws[i] = draws_dev ? draws_dev[i] : nullptr;
  for (int i = 0; i < 4; ++i) a.dec[i] = dec_dev ? dec_dev[i] : nullptr;
  a.nan_flag = h->nan_flag;
}

// reference-order kernel (zz_device.cuh): the move-at-a-time path and ZZ_MOVE_REFERENCE_ORDER sweeps
int launch_sweep(zz_handle *h, uint64_t step, int tune, uint32_t moves, const double *const draws_dev[9], uint8_t *const dec_dev[4],
                 int64_t g_begin, int64_t g_end) {
  if (g_end <= g_begin) return ZZ_OK;
  if (int rc = ensure_caches(h)) return rc;
  SweepArgs a;
  fill_sweep_args(h, a, step, tune, moves, draws_dev, dec_dev, g_begin, g_end);
  k_sweep<<<gene_grid(h, g_end - g_begin, h->C), BLOCK, 0, h->stream>>>(a);
  ZZ_LAUNCHED(h);
  if (moves & (ZZ_MOVE_YG | ZZ_MOVE_SPIKE)) h->lps_valid = false;   // Yg moved without its A(y) being maintained
  h->stats_fresh = false;
  return ZZ_OK;
}

#ifdef ZZ_DEV_R16
#define ZZ_DISPATCH_R(h, KERNEL, grid, ...) KERNEL<16><<<grid, BLOCK, 0, (h)->stream>>>(__VA_ARGS__)
#else
#define ZZ_DISPATCH_R(h, KERNEL, grid, ...)                                                      \
  do {                                                                                           \
    switch ((h)->R) {                                                                            \
      case 2: KERNEL<2><<<grid, BLOCK, 0, (h)->stream>>>(__VA_ARGS__); break;                           \
      case 4: KERNEL<4><<<grid, BLOCK, 0, (h)->stream>>>(__VA_ARGS__); break;                           \
      case 8: KERNEL<8><<<grid, BLOCK, 0, (h)->stream>>>(__VA_ARGS__); break;                           \
      case 16: KERNEL<16><<<grid, BLOCK, 0, (h)->stream>>>(__VA_ARGS__); break;                         \
      default: KERNEL<0><<<grid, BLOCK, 0, (h)->stream>>>(__VA_ARGS__); break;                          \
    }                                                                                            \
  } while (0)
#endif

void fill_fast_args(zz_handle *h, FastArgs &fa) {
  fa.lps = h->lps; fa.exp_tab = h->exp_tab; fa.log_tab = h->log_tab;
  fa.nd_mask = h->nd_mask; fa.xbar = h->xbar; fa.xss = h->xss; fa.fasth = h->fasth;
  fa.partials = h->sw_partials; fa.rows_per_chain = 0; fa.want_stats = 0; fa.do_spike_move = 0;
  fa.done_counter = h->tile_counter; fa.stats_out = nullptr; fa.do_allreduce = 0; memset(&fa.cm, 0, sizeof fa.cm);
  fa.slot_list = nullptr; fa.host_io = 0; memset(&fa.hio, 0, sizeof fa.hio); fa.gate = nullptr; fa.gate_stride = 0;
}

int ensure_fasth(zz_handle *h) {
  if (h->fasth_valid) return ZZ_OK;
  k_prepare_hyper<<<(h->C + 63) / 64, 64, 0, h->stream>>>(h->C, h->hyper, h->fasth);
  ZZ_LAUNCHED(h);
  h->fasth_valid = true;
  return ZZ_OK;
}

int ensure_lps(zz_handle *h, bool rebuild_allzero_xglik = false) {
  if (int rc = ensure_fasth(h)) return rc;
  if (h->lps_valid) return ZZ_OK;
  FastArgs fa;
  fill_sweep_args(h, fa.s, 0, 0, 0, nullptr, nullptr, 0, h->G);
  fill_fast_args(h, fa);
  fa.want_stats = rebuild_allzero_xglik ? 1 : 0;
  ZZ_DISPATCH_R(h, k_refresh_lps, gene_grid(h, h->G, h->C), fa);
  ZZ_LAUNCHED(h);
  h->lps_valid = true;
  return ZZ_OK;
}

// Programmatic dependent launch: the sweep kernels announce their dependents at once (griddepcontrol.launch_dependents) and
// touch mutable state only after griddepcontrol.wait, so the next sweep's CTAs are scheduled as this one's exit and load their
// tables while the last tiles and the statistics epilogue of this sweep are still running.
template <class Kern>
void launch_pdl(Kern kern, dim3 grid, size_t smem, cudaStream_t st, const FastArgs &fa) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(SWEEP_BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("ZZ_NO_PDL") != nullptr;   // measurement switch (profiles/r1_summary.md)
  cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kern, fa);
}

#ifdef ZZ_DEV_R16   /* kernel-tuning builds (tools/): only the R = 16 instantiations, a fifth of the compile time */
#define ZZ_SWEEP_CASE_R(h, PRODV, KTV, grid, smem, args) launch_pdl(k_sweep_fast<16, PRODV, KTV>, grid, smem, (h)->stream, args);
#else
#define ZZ_SWEEP_CASE_R(h, PRODV, KTV, grid, smem, args)                                         \
  switch ((h)->R) {                                                                              \
    case 2: launch_pdl(k_sweep_fast<2, PRODV, KTV>, grid, smem, (h)->stream, args); break;   \
    case 4: launch_pdl(k_sweep_fast<4, PRODV, KTV>, grid, smem, (h)->stream, args); break;   \
    case 8: launch_pdl(k_sweep_fast<8, PRODV, KTV>, grid, smem, (h)->stream, args); break;   \
    case 16: launch_pdl(k_sweep_fast<16, PRODV, KTV>, grid, smem, (h)->stream, args); break; \
    default: launch_pdl(k_sweep_fast<0, PRODV, KTV>, grid, smem, (h)->stream, args); break;  \
  }
#endif
// kt: 0 = generic tile (explicit draws / decisions / move subsets / finite variance bound / K > 3), 1..3 = straight-line tile
#define ZZ_DISPATCH_SWEEP(h, prod, kt, grid, smem, args)                                         \
  do {                                                                                           \
    if ((kt) == 1) { ZZ_SWEEP_CASE_R(h, true, 1, grid, smem, args) }                             \
    else if ((kt) == 2) { ZZ_SWEEP_CASE_R(h, true, 2, grid, smem, args) }                        \
    else if ((kt) == 3) { ZZ_SWEEP_CASE_R(h, true, 3, grid, smem, args) }                        \
    else if (prod) { ZZ_SWEEP_CASE_R(h, true, 0, grid, smem, args) }                             \
    else { ZZ_SWEEP_CASE_R(h, false, 0, grid, smem, args) }                                      \
  } while (0)

template <int RT> void allow_big_smem(int bytes) {
  cudaFuncSetAttribute(k_sweep_fast<RT, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k_sweep_fast<RT, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k_sweep_fast<RT, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k_sweep_fast<RT, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(k_sweep_fast<RT, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// resident CTAs of the persistent sweep kernel on this device (blocks per SM from the occupancy calculator x SM count)
int persistent_blocks(zz_handle *h, size_t smem) {
  if (h->persist_blocks > 0) return h->persist_blocks;
  int per_sm = 0, sms = 0;
  cudaError_t e;
#ifdef ZZ_DEV_R16
  allow_big_smem<16>((int)smem);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<16, true, 0>, SWEEP_BLOCK, smem);
#else
  {   // static tables (11.6 KB) + statistic columns exceed the default 48 KB shared-memory limit: opt in explicitly
    switch (h->R) {
      case 2: allow_big_smem<2>((int)smem); break;
      case 4: allow_big_smem<4>((int)smem); break;
      case 8: allow_big_smem<8>((int)smem); break;
      case 16: allow_big_smem<16>((int)smem); break;
      default: allow_big_smem<0>((int)smem); break;
    }
  }
  switch (h->R) {
    case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<2, true, 0>, SWEEP_BLOCK, smem); break;
    case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<4, true, 0>, SWEEP_BLOCK, smem); break;
    case 8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<8, true, 0>, SWEEP_BLOCK, smem); break;
    case 16: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<16, true, 0>, SWEEP_BLOCK, smem); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_fast<0, true, 0>, SWEEP_BLOCK, smem); break;
  }
#endif
  if (e != cudaSuccess || per_sm < 1) per_sm = 2;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device) != cudaSuccess || sms < 1) sms = 148;
  h->persist_blocks = per_sm * sms;
  return h->persist_blocks;
}

// production kernels (zz_fast.cuh): persistent main sweep + spike move over the all-zero gene list; when the whole shard is
// swept they also leave per-block sufficient statistics in sw_partials (consumed by zz_suffstats[_dev] without another pass)
int launch_sweep_fast(zz_handle *h, uint64_t step, int tune, uint32_t moves, const double *const draws_dev[9], uint8_t *const dec_dev[4],
                      int64_t g_begin, int64_t g_end, const int32_t *slot_list = nullptr, const HostChunk *host_io = nullptr) {
  if (g_end <= g_begin) return ZZ_OK;
  if (int rc = ensure_caches(h)) return rc;
  if (!slot_list) { if (int rc = ensure_lps(h)) return rc; }           // list mode: the caller (k_host_in) has just rebuilt lps
  FastArgs fa;
  fill_sweep_args(h, fa.s, step, tune, moves, draws_dev, dec_dev, g_begin, g_end);
  fill_fast_args(h, fa);
  fa.slot_list = slot_list;
  if (host_io) { fa.host_io = 1; fa.hio = *host_io; }
  const bool full = (g_begin == 0 && g_end == h->G) && !slot_list;
  const int nstat = zz_suffstats_len(h->K);
  size_t smem = full ? sizeof(double) * nstat * SWEEP_BLOCK : 0;
  const int64_t ntiles = (g_end - g_begin + 31) / 32;
  int64_t per_chain = (persistent_blocks(h, sizeof(double) * (nstat + 10) * SWEEP_BLOCK) + h->C - 1) / h->C;
  const int64_t need = (ntiles + SWEEP_BLOCK / 32 - 1) / (SWEEP_BLOCK / 32);
  if (per_chain > need) per_chain = need;
  if (per_chain < 1) per_chain = 1;
  fa.want_stats = full ? 1 : 0;
  fa.rows_per_chain = per_chain;
  fa.do_spike_move = (moves & ZZ_MOVE_SPIKE) ? 1 : 0;
  h->stats_fresh = false;
  if (full && h->stats_target && (size_t)nstat * h->C <= (size_t)MAX_NSTAT * 8) {
    fa.stats_out = h->stats_target;
    if (h->stats_allreduce && h->comm_ready) {
      fa.do_allreduce = 1;
      fill_comm(h, fa.cm, nstat * h->C, comm_take_epochs(h, 1) + 1ull);
    }
  }
  const bool prod = !draws_dev[0] && !draws_dev[2] && !draws_dev[4] && !draws_dev[6] && !dec_dev[0] && !dec_dev[1] && !dec_dev[2] && !dec_dev[3];
  if (dec_dev[3] && (moves & ZZ_MOVE_SPIKE)) ZZ_CUDA(h, cudaMemsetAsync(dec_dev[3], 0, (size_t)h->C * h->G, h->stream));
  int kt = 0;
  if (prod && (moves & ~ZZ_MOVE_REFERENCE_ORDER) == ZZ_MOVE_FUSED && h->K <= 3) {
    kt = h->K;
    if (h->R != 2 && h->R != 4 && h->R != 8 && h->R != 16) kt = 0;     // detect_logsum_sl is unrolled over the library count
    for (int c = 0; c < h->C; ++c) {
      const zz_hyper &hy = h->hyper_host[c];
      if (hy.variance_g_upper_bound != INFINITY) kt = 0;
      double amin = INFINITY, amax = 0;
      for (int r = 0; r < h->R; ++r) { amin = fmin(amin, hy.alpha_r[r]); amax = fmax(amax, hy.alpha_r[r]); }
      if (!(amin > 0 && amax <= 16 * amin)) kt = 0;                    // FastH::uclamp_ok: single clamp of t in detect_logsum_sl
    }
  }
#ifdef ZZ_NO_STRAIGHT_TILE
  kt = 0;
#endif
  if (slot_list) kt = 0;                                            // scattered slots: generic tile, chain 0 only
  if (kt > 0) smem = sizeof(double) * (nstat + 10) * SWEEP_BLOCK;   // statistic columns + per-warp staging slots
  ZZ_DISPATCH_SWEEP(h, prod, kt, dim3((unsigned)per_chain, (unsigned)(slot_list ? 1 : h->C)), smem, fa);
  ZZ_LAUNCHED(h);
  if (full) { h->stats_fresh = true; h->sw_rows = fa.rows_per_chain; }
  return ZZ_OK;
}

// shared implementation of the five move entry points and zz_sweep
int run_moves(zz_handle *h, uint64_t step, int tune, uint32_t moves, const double *draws, const int *draw_slots, int ndraw,
              uint8_t *decisions, const int *dec_slots, int ndec) {
  ZZ_CHECK(h);
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  const int64_t CG = (int64_t)h->C * h->G;
  const double *dd[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t *de[4] = {nullptr, nullptr, nullptr, nullptr};
  if (draws) {
    if (int rc = up_g(h, h->scratch, draws, h->G, ndraw * h->C)) return rc;
    for (int j = 0; j < ndraw; ++j) dd[draw_slots[j]] = h->scratch + (int64_t)j * CG;
  }
  if (decisions) for (int j = 0; j < ndec; ++j) de[dec_slots[j]] = h->dec + (int64_t)j * CG;
  const bool reference_order = (moves & ZZ_MOVE_REFERENCE_ORDER) || (moves & ZZ_MOVE_ALLOC_WITHIN);
  int rc = reference_order ? launch_sweep(h, step, tune, moves & ~ZZ_MOVE_REFERENCE_ORDER, dd, de, 0, h->G)
                           : launch_sweep_fast(h, step, tune, moves, dd, de, 0, h->G);
  if (rc) return rc;
  if (decisions) {
    if (int rc2 = down_g(h, decisions, h->dec, h->G, ndec * h->C)) return rc2;
  } else if (draws) {
    ZZ_CUDA(h, cudaStreamSynchronize(h->stream));  // the caller may free `draws` on return
  }
  return ZZ_OK;
}

int reduce_to_host(zz_handle *h, RedArgs &a, int ny, double *out, int nout_per_y) {
  const int nblk = red_blocks(h->G);
  a.partials = h->partials;
  k_reduce<<<dim3(nblk, ny), BLOCK, 0, h->stream>>>(a);
  ZZ_LAUNCHED(h);
  k_reduce_final<<<ny, BLOCK, 0, h->stream>>>(h->partials, nblk, 2, h->red_out);
  ZZ_LAUNCHED(h);
  ZZ_CUDA(h, cudaMemcpyAsync(h->pinned, h->red_out, sizeof(double) * 2 * ny, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int y = 0; y < ny; ++y)
    for (int j = 0; j < nout_per_y; ++j) out[y * nout_per_y + j] = h->pinned[y * 2 + j];
  return ZZ_OK;
}

void fill_red(zz_handle *h, RedArgs &a, int chain, int op) {
  const int64_t off = (int64_t)chain * h->G;
  memset(&a, 0, sizeof a);
  a.G = h->G; a.op = op; a.chain = chain;
  a.Xg = h->Xg; a.gl = h->gl; a.Yg = h->Yg + off; a.vg = h->vg + off; a.xglik = h->xglik + off; a.vgprob = h->vgprob + off;
  a.flags = h->flags + off; a.hyper = h->hyper + chain;
}

int launch_eval(zz_handle *h, int which, int chain /* -1 = all chains */, double *out0, double *out1, const int *gate = nullptr) {
  EvalArgs a;
  const int64_t off = chain < 0 ? 0 : (int64_t)chain * h->G;
  a.G = h->G; a.which = which; a.Xg = h->Xg; a.gl = h->gl; a.Yg = h->Yg + off; a.vg = h->vg + off; a.flags = h->flags + off;
  a.hyper = h->hyper + (chain < 0 ? 0 : chain); a.out0 = out0; a.out1 = out1; a.gate = gate; a.gate_stride = 0;
  k_eval<<<gene_grid(h, h->G, chain < 0 ? h->C : 1), BLOCK, 0, h->stream>>>(a);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

// Device-resident segments (k_generations) do not maintain the reference's two per-gene caches; whoever reads them next has them
// rebuilt from the state in the reference's operation order (I:473, I:504)
int ensure_caches(zz_handle *h) {
  if (h->caches_valid) return ZZ_OK;
  h->caches_valid = true;
  return launch_eval(h, EVAL_REFRESH, -1, h->xglik, h->vgprob);
}

int eval_to_host(zz_handle *h, int which, int chain, double *out) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!out) return fail(h, ZZ_ERR_INVALID, "out is NULL");
  if (int rc = launch_eval(h, which, chain, h->scratch, nullptr)) return rc;
  return down_g(h, out, h->scratch, h->G);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI

extern "C" {

int zz_create(const zz_config *cfg, zz_handle **out) {
  if (!cfg || !out) return ZZ_ERR_INVALID;
  *out = nullptr;
  if (cfg->num_genes < 1 || cfg->num_libraries < 2 || cfg->num_libraries > ZZ_MAX_LIBS || cfg->num_active_components < 1 ||
      cfg->num_active_components > ZZ_MAX_COMP || cfg->num_chains < 1 || cfg->num_chains > (1 << 20) ||
      cfg->gene_offset < 0 || cfg->gene_offset + cfg->num_genes > 0xFFFFFF00LL)   /* gene ids 0xFFFFFF00.. are the scalar sub-streams (zz_chain.cuh) */
    return ZZ_ERR_INVALID;
  zz_handle *h = new (std::nothrow) zz_handle();
  if (!h) return ZZ_ERR_NOMEM;
  h->cfg = *cfg; h->G = cfg->num_genes; h->C = cfg->num_chains; h->R = cfg->num_libraries; h->K = cfg->num_active_components;
  h->sticky = false; h->launches = 0; h->have_data = false; h->own_stream = false; h->stream = nullptr;
  h->inv_perm = nullptr; h->pipe_ready = false; h->s_up = h->s_down = nullptr;
  h->pp_buf = nullptr; h->pp_nb = 0;
  h->sink_on = false; h->cand_slots = nullptr; h->sink_prow = h->sink_yrow = 0;
  h->hyper_host_stale = false;
  h->devchain = nullptr; h->chain_stats = nullptr; h->chain_ready = false; h->chain_gen = 0; h->chain_samples = 0; h->chain_step = 0;
  h->dlt = nullptr; h->genctl = nullptr; h->genacc = nullptr; h->gen_blocks = 0; h->gen_blocks_l = 0; h->gen_blocks_c = 0; h->bounce = nullptr; h->bounce_cap = 0; h->bounce_ev_ok = false; h->caches_valid = true;
  *out = h;  // returned even on failure so the caller can read zz_last_error and must zz_destroy
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return fail(h, ZZ_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(h, ZZ_ERR_INVALID, "device ordinal out of range");
  ZZ_CUDA(h, cudaSetDevice(cfg->device));
  ZZ_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  const int64_t G = h->G, CG = (int64_t)h->C * G;
  const int nstat = 3 * (h->K + 1) + NSTAT_FIXED;
  ZZ_CUDA(h, cudaMalloc(&h->Xg, sizeof(double) * G * h->R));
  ZZ_CUDA(h, cudaMalloc(&h->gl, sizeof(double) * G));
  double **d8[] = {&h->Yg, &h->vg, &h->xglik, &h->vgprob, &h->ty, &h->tv};
  for (auto p : d8) { ZZ_CUDA(h, cudaMalloc(p, sizeof(double) * CG)); ZZ_CUDA(h, cudaMemsetAsync(*p, 0, sizeof(double) * CG, h->stream)); }
  ZZ_CUDA(h, cudaMalloc(&h->flags, sizeof(uint16_t) * CG));
  ZZ_CUDA(h, cudaMemsetAsync(h->flags, 0, sizeof(uint16_t) * CG, h->stream));
  ZZ_CUDA(h, cudaMalloc(&h->win_y, sizeof(ulonglong2) * CG));
  ZZ_CUDA(h, cudaMalloc(&h->win_v, sizeof(ulonglong2) * CG));
  ZZ_CUDA(h, cudaMalloc(&h->trace, sizeof(uint32_t) * CG * (h->K + 1)));
  ZZ_CUDA(h, cudaMemsetAsync(h->trace, 0, sizeof(uint32_t) * CG * (h->K + 1), h->stream));
  ZZ_CUDA(h, cudaMalloc(&h->hyper, sizeof(zz_hyper) * h->C));
  ZZ_CUDA(h, cudaMemsetAsync(h->hyper, 0, sizeof(zz_hyper) * h->C, h->stream));
  const size_t npart = (size_t)RED_BLOCKS * (size_t)(h->C > ZZ_MAX_LIBS ? h->C : ZZ_MAX_LIBS) * MAX_NSTAT;
  ZZ_CUDA(h, cudaMalloc(&h->partials, sizeof(double) * npart));
  ZZ_CUDA(h, cudaMalloc(&h->red_out, sizeof(double) * (size_t)(h->C > ZZ_MAX_LIBS ? h->C : ZZ_MAX_LIBS) * MAX_NSTAT));
  ZZ_CUDA(h, cudaMalloc(&h->scratch, sizeof(double) * 9 * CG));
  ZZ_CUDA(h, cudaMalloc(&h->dec, (size_t)4 * CG));
  ZZ_CUDA(h, cudaMalloc(&h->iscratch, sizeof(int32_t) * 3 * CG));
  ZZ_CUDA(h, cudaMalloc(&h->lps, sizeof(double) * CG));
  ZZ_CUDA(h, cudaMemsetAsync(h->lps, 0, sizeof(double) * CG, h->stream));
  ZZ_CUDA(h, cudaMalloc(&h->exp_tab, sizeof(kExpTab)));
  ZZ_CUDA(h, cudaMalloc(&h->log_tab, sizeof(kLogTab)));
  ZZ_CUDA(h, cudaMemcpyAsync(h->exp_tab, kExpTab, sizeof(kExpTab), cudaMemcpyHostToDevice, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(h->log_tab, kLogTab, sizeof(kLogTab), cudaMemcpyHostToDevice, h->stream));
  ZZ_CUDA(h, cudaMalloc(&h->perm, sizeof(int32_t) * G));
  ZZ_CUDA(h, cudaMalloc(&h->stage, sizeof(double) * 9 * CG));
  h->perm_h.resize(G);
  for (int64_t g = 0; g < G; ++g) h->perm_h[g] = (int32_t)g;
  ZZ_CUDA(h, cudaMemcpyAsync(h->perm, h->perm_h.data(), sizeof(int32_t) * G, cudaMemcpyHostToDevice, h->stream));
  h->stats_target = nullptr; h->stats_allreduce = 0;
  h->comm_rank = 0; h->comm_world = 1; h->comm_epoch = 0; h->comm_epoch_small = 0; h->comm_ready = false;
  h->comm_timeout_ns = (getenv("ZZ_COMM_TIMEOUT_S") ? atoll(getenv("ZZ_COMM_TIMEOUT_S")) : 20) * 1000000000LL; h->comm_inbox = nullptr; h->comm_flags = nullptr;
  for (int r = 0; r < 8; ++r) { h->peer_inbox[r] = nullptr; h->peer_flags[r] = nullptr; }
  h->n_allzero = 0; h->lps_valid = false; h->stats_fresh = false; h->persist_blocks = 0; h->sw_rows = 0;
  ZZ_CUDA(h, cudaMalloc(&h->fasth, sizeof(FastH) * h->C));
  h->fasth_valid = false;
  ZZ_CUDA(h, cudaMalloc(&h->nd_mask, sizeof(uint64_t) * G));
  ZZ_CUDA(h, cudaMalloc(&h->xbar, sizeof(double) * G));
  ZZ_CUDA(h, cudaMalloc(&h->xss, sizeof(double) * G));
  ZZ_CUDA(h, cudaMalloc(&h->tile_counter, sizeof(unsigned) * 32));
  ZZ_CUDA(h, cudaMemsetAsync(h->tile_counter, 0, sizeof(unsigned) * 32, h->stream));
  {
    const int64_t max_rows = (G + 31) / 32 + ((G + BLOCK - 1) / BLOCK) * (BLOCK / 32);
    ZZ_CUDA(h, cudaMalloc(&h->sw_partials, sizeof(double) * (size_t)h->C * max_rows * nstat));
  }
  ZZ_CUDA(h, cudaMalloc(&h->nan_flag, sizeof(int)));
  ZZ_CUDA(h, cudaMemsetAsync(h->nan_flag, 0, sizeof(int), h->stream));
  ZZ_CUDA(h, cudaMallocHost(&h->pinned, sizeof(double) * (size_t)(h->C > ZZ_MAX_LIBS ? h->C : ZZ_MAX_LIBS) * MAX_NSTAT));
  h->hyper_host.assign(h->C, zz_hyper());
  k_window_init<<<gene_grid(h, CG), BLOCK, 0, h->stream>>>(CG, h->win_y, h->win_v);
  ZZ_LAUNCHED(h);
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

void zz_destroy(zz_handle *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (int r = 0; r < 8; ++r) {
    if (r != h->comm_rank && h->peer_inbox[r]) cudaIpcCloseMemHandle(h->peer_inbox[r]);
    if (r != h->comm_rank && h->peer_flags[r]) cudaIpcCloseMemHandle(h->peer_flags[r]);
  }
  if (h->comm_inbox) cudaFree(h->comm_inbox);
  if (h->comm_flags) cudaFree(h->comm_flags);
  void *ptrs[] = {h->Xg, h->gl, h->Yg, h->vg, h->xglik, h->vgprob, h->ty, h->tv, h->flags, h->win_y, h->win_v, h->trace,
                  h->hyper, h->partials, h->red_out, h->scratch, h->dec, h->iscratch, h->nan_flag, h->lps, h->exp_tab,
                  h->log_tab, h->perm, h->stage, h->nd_mask, h->xbar, h->xss, h->tile_counter, h->sw_partials, h->fasth};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->bounce) cudaFreeHost(h->bounce);
  if (h->bounce_ev_ok) for (cudaEvent_t e : h->bounce_ev) cudaEventDestroy(e);
  if (h->inv_perm) cudaFree(h->inv_perm);
  if (h->devchain) cudaFree(h->devchain);
  if (h->cand_slots) cudaFree(h->cand_slots);
  if (h->pp_buf) cudaFree(h->pp_buf);
  if (h->chain_stats) cudaFree(h->chain_stats);
  if (h->dlt) cudaFree(h->dlt);
  if (h->genctl) cudaFree(h->genctl);
  if (h->genacc) cudaFree(h->genacc);
  if (h->pipe_ready) {
    cudaStreamDestroy(h->s_up); cudaStreamDestroy(h->s_down); cudaEventDestroy(h->ev_entry);
    for (int k = 0; k < ZZ_HOST_MAX_CHUNKS; ++k) { cudaEventDestroy(h->ev_up[k]); cudaEventDestroy(h->ev_done[k]); }
  }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *zz_last_error(const zz_handle *h) { return h ? h->err.c_str() : "null handle"; }

int zz_set_stream(zz_handle *h, void *cuda_stream) {
  ZZ_CHECK(h);
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = (cudaStream_t)cuda_stream;
  return ZZ_OK;
}

int zz_sync(zz_handle *h) {
  ZZ_CHECK(h);
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

uint64_t zz_launch_count(const zz_handle *h) { return h ? h->launches : 0; }

int zz_upload_data(zz_handle *h, const double *Xg, const double *gene_lengths) {
  ZZ_CHECK(h);
  if (!Xg || !gene_lengths) return fail(h, ZZ_ERR_INVALID, "Xg / gene_lengths is NULL");
  if (h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data may be called once per handle (the gene order on the device is fixed by it)");
  const int64_t G = h->G;
  const int R = h->R;
  // Device gene order: genes with at least one read sorted by decreasing (mean detected log-expression + log gene length), then the all-zero
  // genes (no_detect_spike, I:416-417).  Warps therefore hold genes of similar expression: the detection term saturates
  // (p_x == 1) for whole warps of highly expressed genes, and the spike-allocation candidates are contiguous.
  std::vector<double> key(G);
  std::vector<uint8_t> allzero(G);
  for (int64_t g = 0; g < G; ++g) {
    double sum = 0; int n = 0;
    for (int r = 0; r < R; ++r) { const double x = Xg[(int64_t)r * G + g]; if (x != -INFINITY) { sum += x; n++; } }
    allzero[g] = (n == 0);
    key[g] = n ? sum / n + log(gene_lengths[g]) : -INFINITY;   // ~ log(gl * exp(y)): the quantity the detection term saturates in
  }
  std::vector<int32_t> &perm = h->perm_h;
  for (int64_t g = 0; g < G; ++g) perm[g] = (int32_t)g;
  std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) {
    if (allzero[a] != allzero[b]) return allzero[a] < allzero[b];
    return key[a] > key[b];
  });
  int64_t naz = 0;
  for (int64_t g = 0; g < G; ++g) naz += allzero[g];
  h->n_allzero = naz;
  ZZ_CUDA(h, cudaMemcpyAsync(h->perm, perm.data(), sizeof(int32_t) * G, cudaMemcpyHostToDevice, h->stream));
  {
    std::vector<double> tmp((size_t)G * (R + 1));
    for (int r = 0; r < R; ++r)
      for (int64_t j = 0; j < G; ++j) tmp[(size_t)r * G + j] = Xg[(int64_t)r * G + perm[j]];
    for (int64_t j = 0; j < G; ++j) tmp[(size_t)R * G + j] = gene_lengths[perm[j]];
    ZZ_CUDA(h, cudaMemcpyAsync(h->Xg, tmp.data(), sizeof(double) * G * R, cudaMemcpyHostToDevice, h->stream));
    ZZ_CUDA(h, cudaMemcpyAsync(h->gl, tmp.data() + (size_t)R * G, sizeof(double) * G, cudaMemcpyHostToDevice, h->stream));
    std::vector<uint16_t> f((size_t)h->C * G);
    for (int c = 0; c < h->C; ++c)
      for (int64_t j = 0; j < G; ++j) f[(size_t)c * G + j] = allzero[perm[j]] ? (uint16_t)F_ALLZERO : (uint16_t)0;
    ZZ_CUDA(h, cudaMemcpyAsync(h->flags, f.data(), sizeof(uint16_t) * f.size(), cudaMemcpyHostToDevice, h->stream));
    ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  k_prepare_data<<<gene_grid(h, G), BLOCK, 0, h->stream>>>(G, R, h->Xg, h->nd_mask, h->xbar, h->xss);
  ZZ_LAUNCHED(h);
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  h->have_data = true; h->lps_valid = false; h->stats_fresh = false;
  return ZZ_OK;
}

int zz_set_state(zz_handle *h, int chain, const double *Yg, const double *variance_g, const int32_t *ai, const int32_t *within,
                 const int32_t *spike, const double *XgLikelihood, const double *variance_g_probability, const double *ty,
                 const double *tv, const int32_t *pinned_active) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  const int64_t G = h->G, off = (int64_t)chain * G;
  int rc;
  if (Yg || spike) h->lps_valid = false;
  h->stats_fresh = false;
  if (XgLikelihood && variance_g_probability && h->C == 1) h->caches_valid = true;
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  if ((rc = up_g(h, h->Yg + off, Yg, G)) || (rc = up_g(h, h->vg + off, variance_g, G)) || (rc = up_g(h, h->xglik + off, XgLikelihood, G)) ||
      (rc = up_g(h, h->vgprob + off, variance_g_probability, G)) || (rc = up_g(h, h->ty + off, ty, G)) || (rc = up_g(h, h->tv + off, tv, G)))
    return rc;
  if (ai || within || spike || pinned_active) {
    int32_t *d_ai = h->iscratch, *d_w = h->iscratch + G, *d_s = h->iscratch + 2 * G, *d_p = reinterpret_cast<int32_t *>(h->stage);
    if ((rc = up(h, d_ai, ai, G)) || (rc = up(h, d_w, within, G)) || (rc = up(h, d_s, spike, G)) || (rc = up(h, d_p, pinned_active, G))) return rc;
    k_pack_flags<<<gene_grid(h, G), BLOCK, 0, h->stream>>>(G, h->perm, ai ? d_ai : nullptr, within ? d_w : nullptr, spike ? d_s : nullptr,
                                                         pinned_active ? d_p : nullptr, h->flags + off);
    ZZ_LAUNCHED(h);
  }
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

int zz_get_state(zz_handle *h, int chain, double *Yg, double *variance_g, int32_t *ai, int32_t *within, int32_t *spike,
                 double *XgLikelihood, double *variance_g_probability, double *ty, double *tv) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  const int64_t G = h->G, off = (int64_t)chain * G;
  int rc;
  if ((XgLikelihood || variance_g_probability) && (rc = ensure_caches(h))) return rc;
  // Every requested array is scattered into the caller's gene order in its own part of the device staging buffer and copied down
  // asynchronously.  A destination in pinned memory receives the copy directly.  A PAGEABLE destination (an R vector, a NumPy
  // array) would make the driver stage the copy at ~6 GB/s: those land in a pinned bounce region of the handle at PCIe speed and are
  // copied out (and their fresh pages faulted in) by four host threads while the following arrays are still in flight.
  struct Out { void *dst; const void *dev; size_t bytes, boff; bool direct; };
  Out outs[9];
  int n_out = 0;
  size_t need = 0;
  double *st = h->stage;                                             // 9 * C * G doubles
  const double *src8[6] = {h->Yg + off, h->vg + off, h->xglik + off, h->vgprob + off, h->ty + off, h->tv + off};
  double *dst8[6] = {Yg, variance_g, XgLikelihood, variance_g_probability, ty, tv};
  auto is_pinned = [](const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice;   // a device pointer is handed to the copy as it is (and refused there), not to memcpy
  };
  for (int j = 0; j < 6; ++j) {
    if (!dst8[j]) continue;
    double *part = st + (int64_t)j * G;
    k_scatter<double><<<dim3((unsigned)((G + BLOCK - 1) / BLOCK), 1), BLOCK, 0, h->stream>>>(G, h->perm, src8[j], part);
    ZZ_LAUNCHED(h);
    outs[n_out++] = Out{dst8[j], part, sizeof(double) * (size_t)G, 0, is_pinned(dst8[j])};
  }
  if (ai || within || spike) {
    int32_t *d_ai = h->iscratch, *d_w = h->iscratch + G, *d_s = h->iscratch + 2 * G;
    k_unpack_flags<<<gene_grid(h, G), BLOCK, 0, h->stream>>>(G, h->perm, h->flags + off, d_ai, d_w, d_s);
    ZZ_LAUNCHED(h);
    int32_t *dsti[3] = {ai, within, spike}, *srci[3] = {d_ai, d_w, d_s};
    for (int j = 0; j < 3; ++j) if (dsti[j]) outs[n_out++] = Out{dsti[j], srci[j], sizeof(int32_t) * (size_t)G, 0, is_pinned(dsti[j])};
  }
  for (int j = 0; j < n_out; ++j) if (!outs[j].direct) { outs[j].boff = need; need += (outs[j].bytes + 255) & ~(size_t)255; }
  if (need > h->bounce_cap) {
    if (h->bounce) { ZZ_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFreeHost(h->bounce); h->bounce = nullptr; h->bounce_cap = 0; }
    ZZ_CUDA(h, cudaHostAlloc(reinterpret_cast<void **>(&h->bounce), need, cudaHostAllocDefault));
    h->bounce_cap = need;
  }
  if (need && !h->bounce_ev_ok) {
    for (cudaEvent_t &e : h->bounce_ev) ZZ_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->bounce_ev_ok = true;
  }
  for (int j = 0; j < n_out; ++j) {
    void *to = outs[j].direct ? outs[j].dst : static_cast<void *>(h->bounce + outs[j].boff);
    ZZ_CUDA(h, cudaMemcpyAsync(to, outs[j].dev, outs[j].bytes, cudaMemcpyDeviceToHost, h->stream));
    if (!outs[j].direct) ZZ_CUDA(h, cudaEventRecord(h->bounce_ev[j], h->stream));
  }
  for (int j = 0; j < n_out; ++j) {
    if (outs[j].direct) continue;
    ZZ_CUDA(h, cudaEventSynchronize(h->bounce_ev[j]));
    const char *from = h->bounce + outs[j].boff;
    char *to = static_cast<char *>(outs[j].dst);
    const size_t n = outs[j].bytes;
    static const int nt_env = getenv("ZZ_HOST_COPY_THREADS") ? atoi(getenv("ZZ_HOST_COPY_THREADS")) : 4;
    const int nt = n >= ((size_t)2 << 20) ? std::max(1, std::min(nt_env, 16)) : 1;   // host threads per array
    if (nt == 1) { memcpy(to, from, n); continue; }
    const size_t per = ((n / nt) + 4095) & ~(size_t)4095;
    std::thread th[15];
    size_t done_to = std::min(n, per);                              // this thread's own part; grows if a helper cannot be started
    for (int t = 1; t < nt; ++t) {
      const size_t b = per * t, e = std::min(n, per * (t + 1));
      if (b >= e) break;
      try { th[t - 1] = std::thread([=] { memcpy(to + b, from + b, e - b); }); }
      catch (...) { memcpy(to + b, from + b, e - b); }               // no thread to be had (resource limits): copy that part here
    }
    memcpy(to, from, done_to);
    for (std::thread &t : th) if (t.joinable()) t.join();
  }
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

int zz_set_hyper(zz_handle *h, int chain, const zz_hyper *hy) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!hy) return fail(h, ZZ_ERR_INVALID, "hyper is NULL");
  if (hy->num_libraries != h->R || hy->num_active_components != h->K)
    return fail(h, ZZ_ERR_INVALID, "hyper dimensions do not match the handle");
  if (h->hyper_host_stale || memcmp(h->hyper_host[chain].alpha_r, hy->alpha_r, sizeof(double) * h->R) != 0) h->lps_valid = false;  // A depends on alpha_r
  h->hyper_host_stale = false;
  h->fasth_valid = false;
  h->hyper_host[chain] = *hy;
  ZZ_CUDA(h, cudaMemcpyAsync(h->hyper + chain, &h->hyper_host[chain], sizeof(zz_hyper), cudaMemcpyHostToDevice, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

int zz_get_hyper(zz_handle *h, int chain, zz_hyper *hy) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!hy) return fail(h, ZZ_ERR_INVALID, "hyper is NULL");
  ZZ_CUDA(h, cudaMemcpyAsync(hy, h->hyper + chain, sizeof(zz_hyper), cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

int zz_refresh_caches(zz_handle *h) {
  ZZ_CHECK(h);
  h->stats_fresh = false;
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  h->caches_valid = true;
  return launch_eval(h, EVAL_REFRESH, -1, h->xglik, h->vgprob);
}

int zz_debug_refresh_lps(zz_handle *h) {
  ZZ_CHECK(h);
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  h->lps_valid = false;
  return ensure_lps(h);
}

int zz_reset_windows(zz_handle *h) {
  ZZ_CHECK(h);
  const int64_t CG = (int64_t)h->C * h->G;
  k_window_init<<<gene_grid(h, CG), BLOCK, 0, h->stream>>>(CG, h->win_y, h->win_v);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

int zz_get_window_sums(zz_handle *h, int chain, int32_t *yg_sum, int32_t *vg_sum) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  const int64_t G = h->G, off = (int64_t)chain * G;
  k_window_sums<<<gene_grid(h, G), BLOCK, 0, h->stream>>>(G, h->win_y + off, h->win_v + off, h->iscratch, h->iscratch + G);
  ZZ_LAUNCHED(h);
  int rc;
  if ((rc = down_g(h, yg_sum, h->iscratch, G)) || (rc = down_g(h, vg_sum, h->iscratch + G, G))) return rc;
  return ZZ_OK;
}

int zz_eval_xg_loglik(zz_handle *h, int chain, double *out) { return eval_to_host(h, EVAL_XGLIK, chain, out); }
int zz_eval_varg_logprior(zz_handle *h, int chain, double *out) { return eval_to_host(h, EVAL_VGPRIOR, chain, out); }
int zz_eval_level2_loglik(zz_handle *h, int chain, double *out) { return eval_to_host(h, EVAL_LEVEL2, chain, out); }

int zz_mh_yg(zz_handle *h, uint64_t step, int tune, const double *draws, uint8_t *decisions) {
  static const int ds[] = {0, 1}, es[] = {0};
  return run_moves(h, step, tune, ZZ_MOVE_YG | ZZ_MOVE_REFERENCE_ORDER, draws, ds, 2, decisions, es, 1);
}
int zz_mh_variance_g(zz_handle *h, uint64_t step, int tune, const double *draws, uint8_t *decisions) {
  static const int ds[] = {2, 3}, es[] = {1};
  return run_moves(h, step, tune, ZZ_MOVE_VARIANCE_G | ZZ_MOVE_REFERENCE_ORDER, draws, ds, 2, decisions, es, 1);
}
int zz_gibbs_alloc(zz_handle *h, uint64_t step, const double *draws, uint8_t *decisions) {
  static const int ds[] = {4, 5}, es[] = {2};
  return run_moves(h, step, 0, ZZ_MOVE_ALLOC | ZZ_MOVE_REFERENCE_ORDER, draws, ds, 2, decisions, es, 1);
}
int zz_gibbs_alloc_within(zz_handle *h, uint64_t step, const double *draws, uint8_t *decisions) {
  static const int ds[] = {5}, es[] = {2};
  return run_moves(h, step, 0, ZZ_MOVE_ALLOC_WITHIN | ZZ_MOVE_REFERENCE_ORDER, draws, ds, 1, decisions, es, 1);
}
int zz_mh_spike_alloc(zz_handle *h, uint64_t step, const double *draws, uint8_t *decisions) {
  static const int ds[] = {6, 7, 8}, es[] = {3};
  return run_moves(h, step, 0, ZZ_MOVE_SPIKE | ZZ_MOVE_REFERENCE_ORDER, draws, ds, 3, decisions, es, 1);
}
int zz_sweep(zz_handle *h, uint64_t step, int tune, uint32_t moves, const double *draws, uint8_t *decisions) {
  static const int ds[] = {0, 1, 2, 3, 4, 5, 6, 7, 8}, es[] = {0, 1, 2, 3};
  if ((moves & ~ZZ_MOVE_REFERENCE_ORDER) == 0) moves |= ZZ_MOVE_FUSED;
  return run_moves(h, step, tune, moves, draws, ds, 9, decisions, es, 4);
}

int zz_philox_draws(zz_handle *h, uint64_t step, int slot, double *out) {
  ZZ_CHECK(h);
  int n = 0;
  switch (slot) {
    case ZZ_SLOT_YG: case ZZ_SLOT_VARIANCE_G: case ZZ_SLOT_ALLOC: n = 2; break;
    case ZZ_SLOT_ALLOC_WITHIN: n = 1; break;
    case ZZ_SLOT_SPIKE: n = 3; break;
    default: return fail(h, ZZ_ERR_INVALID, "unknown move slot");
  }
  if (!out) return fail(h, ZZ_ERR_INVALID, "out is NULL");
  const int64_t CG = (int64_t)h->C * h->G;
  k_draws<<<gene_grid(h, h->G, h->C), BLOCK, 0, h->stream>>>(h->G, h->cfg.gene_offset, h->perm, h->C, h->cfg.seed, step, slot, h->scratch);
  ZZ_LAUNCHED(h);
  (void)CG;
  return down_g(h, out, h->scratch, h->G, n * h->C);
}

int zz_suffstats_len(int K) { return 3 * (K + 1) + NSTAT_FIXED; }

int zz_suffstats_dev(zz_handle *h, double *out_dev) {
  ZZ_CHECK(h);
  if (!out_dev) return fail(h, ZZ_ERR_INVALID, "out_dev is NULL");
  if (h->stats_fresh) {   // the last production sweep already emitted per-tile statistics: two tiny fixed-order reductions
    k_reduce_final<<<h->C, BLOCK, 0, h->stream>>>(h->sw_partials, (int)h->sw_rows, zz_suffstats_len(h->K), out_dev);
    ZZ_LAUNCHED(h);
    return ZZ_OK;
  }
  const int nblk = red_blocks(h->G), nstat = zz_suffstats_len(h->K);
  if (int rc = ensure_caches(h)) return rc;
  StatArgs a;
  a.G = h->G; a.K = h->K; a.Yg = h->Yg; a.vg = h->vg; a.xglik = h->xglik; a.vgprob = h->vgprob; a.flags = h->flags; a.partials = h->partials;
  k_suffstats<<<dim3(nblk, h->C), BLOCK, 0, h->stream>>>(a);
  ZZ_LAUNCHED(h);
  k_reduce_final<<<h->C, BLOCK, 0, h->stream>>>(h->partials, nblk, nstat, out_dev);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

int zz_suffstats(zz_handle *h, double *out) {
  ZZ_CHECK(h);
  if (!out) return fail(h, ZZ_ERR_INVALID, "out is NULL");
  if (int rc = zz_suffstats_dev(h, h->red_out)) return rc;
  const int n = zz_suffstats_len(h->K) * h->C;
  ZZ_CUDA(h, cudaMemcpyAsync(h->pinned, h->red_out, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  memcpy(out, h->pinned, sizeof(double) * n);
  return ZZ_OK;
}


int zz_comm_handle_bytes(void) { return 2 * (int)sizeof(cudaIpcMemHandle_t); }

int zz_comm_init(zz_handle *h, int rank, int world, void *handle_out) {
  ZZ_CHECK(h);
  if (world < 1 || world > 8 || rank < 0 || rank >= world || !handle_out) return fail(h, ZZ_ERR_INVALID, "zz_comm_init: need 0 <= rank < world <= 8");
  if (h->comm_inbox) return fail(h, ZZ_ERR_STATE, "zz_comm_init was already called on this handle");
  const size_t nvec = (size_t)zz_suffstats_len(h->K) * h->C;
  if (nvec > (size_t)MAX_NSTAT * 8) return fail(h, ZZ_ERR_INVALID, "zz_comm_init: too many chains for the fused all-reduce (nstat * chains <= 296)");
  h->comm_rank = rank; h->comm_world = world; h->comm_epoch = 0; h->comm_epoch_small = 0;
  // [2 slots][world][2 words per double] for the statistics vector, then the same for the small gene sums of zz_run_generations
  ZZ_CUDA(h, cudaMalloc(&h->comm_inbox, sizeof(double) * 2 * world * 2 * (nvec + COMM_SMALL_NVEC)));
  ZZ_CUDA(h, cudaMalloc(&h->comm_flags, sizeof(unsigned long long) * 2 * world));
  ZZ_CUDA(h, cudaMemset(h->comm_inbox, 0, sizeof(double) * 2 * world * 2 * (nvec + COMM_SMALL_NVEC)));
  ZZ_CUDA(h, cudaMemset(h->comm_flags, 0, sizeof(unsigned long long) * 2 * world));
  cudaIpcMemHandle_t hs[2];
  ZZ_CUDA(h, cudaIpcGetMemHandle(&hs[0], h->comm_inbox));
  ZZ_CUDA(h, cudaIpcGetMemHandle(&hs[1], h->comm_flags));
  memcpy(handle_out, hs, sizeof hs);
  return ZZ_OK;
}

int zz_comm_connect(zz_handle *h, const void *all_handles) {
  ZZ_CHECK(h);
  if (!h->comm_inbox || !all_handles) return fail(h, ZZ_ERR_STATE, "zz_comm_connect: call zz_comm_init first");
  const cudaIpcMemHandle_t *hs = reinterpret_cast<const cudaIpcMemHandle_t *>(all_handles);
  for (int r = 0; r < h->comm_world; ++r) {
    if (r == h->comm_rank) { h->peer_inbox[r] = h->comm_inbox; h->peer_flags[r] = h->comm_flags; continue; }
    void *p0 = nullptr, *p1 = nullptr;
    ZZ_CUDA(h, cudaIpcOpenMemHandle(&p0, hs[2 * r + 0], cudaIpcMemLazyEnablePeerAccess));
    ZZ_CUDA(h, cudaIpcOpenMemHandle(&p1, hs[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
    h->peer_inbox[r] = (double *)p0; h->peer_flags[r] = (unsigned long long *)p1;
  }
  h->comm_ready = true;
  return ZZ_OK;
}

int zz_set_stats_output_dev(zz_handle *h, double *out_dev, int allreduce) {
  ZZ_CHECK(h);
  if (allreduce && !h->comm_ready) return fail(h, ZZ_ERR_STATE, "zz_set_stats_output_dev: all-reduce requested before zz_comm_connect");
  // (the epilogue of k_sweep_fast holds at most MAX_NSTAT * 8 values: a handle with more chains gets its statistics from
  //  zz_run_sweeps / zz_run_generations, whose kernel writes them chain by chain, or from zz_suffstats)
  h->stats_target = out_dev; h->stats_allreduce = allreduce ? 1 : 0;
  return ZZ_OK;
}

int zz_suffstats_allreduce_dev(zz_handle *h, double *out_dev) {
  ZZ_CHECK(h);
  if (!out_dev) return fail(h, ZZ_ERR_INVALID, "out_dev is NULL");
  if (!h->comm_ready) return fail(h, ZZ_ERR_STATE, "zz_comm_connect has not been called");
  const int nstat = zz_suffstats_len(h->K);
  const double *partials; int nblk;
  if (h->stats_fresh) { partials = h->sw_partials; nblk = (int)h->sw_rows; }
  else {
    if (int rc = ensure_caches(h)) return rc;
    nblk = red_blocks(h->G);
    StatArgs a;
    a.G = h->G; a.K = h->K; a.Yg = h->Yg; a.vg = h->vg; a.xglik = h->xglik; a.vgprob = h->vgprob; a.flags = h->flags; a.partials = h->partials;
    k_suffstats<<<dim3(nblk, h->C), BLOCK, 0, h->stream>>>(a);
    ZZ_LAUNCHED(h);
    partials = h->partials;
  }
  CommArgs cm;
  fill_comm(h, cm, nstat * h->C, comm_take_epochs(h, 1) + 1ull);
  k_reduce_allreduce<<<1, BLOCK, 0, h->stream>>>(partials, nblk, nstat, h->C, out_dev, cm, h->nan_flag);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

int zz_varg_hyper_sum(zz_handle *h, int chain, double s0, double s1, double tau, double *out2) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!out2) return fail(h, ZZ_ERR_INVALID, "out is NULL");
  if (int rc = ensure_caches(h)) return rc;
  RedArgs a; fill_red(h, a, chain, RED_VARG_HYPER);
  a.s0 = s0; a.s1 = s1; a.tau = tau;
  return reduce_to_host(h, a, 1, out2, 2);
}

int zz_commit_varg_hyper(zz_handle *h, int chain) {
  ZZ_CHECK(h);
  h->stats_fresh = false;
  if (int rc = chain_ok(h, chain)) return rc;
  if (int rc = ensure_caches(h)) return rc;
  return launch_eval(h, EVAL_COMMIT_VG, chain, nullptr, h->vgprob + (int64_t)chain * h->G);
}

int zz_level2_sums(zz_handle *h, int chain, double im, double iv, const double *am, const double *av, double *out2) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!am || !av || !out2) return fail(h, ZZ_ERR_INVALID, "NULL argument");
  RedArgs a; fill_red(h, a, chain, RED_LEVEL2);
  a.im = im; a.iv = iv;
  for (int k = 0; k < h->K; ++k) { a.am[k] = am[k]; a.av[k] = av[k]; }
  for (int k = h->K; k < ZZ_MAX_COMP; ++k) { a.am[k] = 0; a.av[k] = 1; }
  return reduce_to_host(h, a, 1, out2, 2);
}

int zz_px_delta(zz_handle *h, int chain, const double *alpha_prop, const uint8_t *lib_mask, double *out) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!alpha_prop || !out) return fail(h, ZZ_ERR_INVALID, "NULL argument");
  RedArgs a; fill_red(h, a, chain, RED_PXDELTA);
  for (int r = 0; r < h->R; ++r) { a.alpha_prop[r] = alpha_prop[r]; a.lib_mask[r] = lib_mask ? lib_mask[r] : 1; }
  return reduce_to_host(h, a, h->R, out, 1);
}

int zz_px_commit(zz_handle *h, int chain, int lib, double alpha_new) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (lib < 0 || lib >= h->R) return fail(h, ZZ_ERR_INVALID, "library index out of range");
  if (h->hyper_host_stale) {   // a device-resident segment moved the hyper-parameters: this entry uploads the host copy below
    ZZ_CUDA(h, cudaMemcpyAsync(&h->hyper_host[0], h->hyper, sizeof(zz_hyper), cudaMemcpyDeviceToHost, h->stream));
    ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
    h->hyper_host_stale = false;
  }
  if (int rc = ensure_caches(h)) return rc;
  const int64_t off = (int64_t)chain * h->G;
  if (h->R > 2) {  // Pr:213-215
    k_px_commit<<<gene_grid(h, h->G), BLOCK, 0, h->stream>>>(h->G, lib, alpha_new, h->Xg, h->gl, h->Yg + off, h->vg + off,
                                                            h->flags + off, h->hyper + chain, h->xglik + off);
    ZZ_LAUNCHED(h);
  }
  h->hyper_host[chain].alpha_r[lib] = alpha_new;
  h->lps_valid = false; h->stats_fresh = false; h->fasth_valid = false;
  ZZ_CUDA(h, cudaMemcpyAsync(h->hyper + chain, &h->hyper_host[chain], sizeof(zz_hyper), cudaMemcpyHostToDevice, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->R <= 2) {  // Pr:219: full recompute with the proposed p_x
    if (int rc = launch_eval(h, EVAL_XGLIK, chain, h->xglik + off, nullptr)) return rc;
  }
  return ZZ_OK;
}

int zz_lnl(zz_handle *h, int chain, double *out) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!out) return fail(h, ZZ_ERR_INVALID, "out is NULL");
  RedArgs a; fill_red(h, a, chain, RED_LNL);
  return reduce_to_host(h, a, 1, out, 1);
}

int zz_reset_alloc_trace(zz_handle *h) {
  ZZ_CHECK(h);
  ZZ_CUDA(h, cudaMemsetAsync(h->trace, 0, sizeof(uint32_t) * (size_t)h->C * h->G * (h->K + 1), h->stream));
  return ZZ_OK;
}

int zz_accumulate_alloc_trace(zz_handle *h) {
  ZZ_CHECK(h);
  k_alloc_trace<<<gene_grid(h, h->G, h->C), BLOCK, 0, h->stream>>>(h->G, h->K, h->flags, h->trace);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

int zz_accumulate_alloc_trace_from(zz_handle *h, int src_chain, int dst_chain) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, src_chain)) return rc;
  if (int rc = chain_ok(h, dst_chain)) return rc;
  k_alloc_trace<<<gene_grid(h, h->G, 1), BLOCK, 0, h->stream>>>(h->G, h->K, h->flags + (int64_t)src_chain * h->G,
                                                                 h->trace + (int64_t)dst_chain * (h->K + 1) * h->G);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

int zz_copy_chain_tuning(zz_handle *h, int src_chain, int dst_chain) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, src_chain)) return rc;
  if (int rc = chain_ok(h, dst_chain)) return rc;
  if (src_chain == dst_chain) return ZZ_OK;
  const int64_t G = h->G;
  ZZ_CUDA(h, cudaMemcpyAsync(h->ty + dst_chain * G, h->ty + src_chain * G, sizeof(double) * G, cudaMemcpyDeviceToDevice, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(h->tv + dst_chain * G, h->tv + src_chain * G, sizeof(double) * G, cudaMemcpyDeviceToDevice, h->stream));
  return ZZ_OK;
}

int zz_init_state(zz_handle *h, uint64_t step, double rowvar_log_mean, double rowvar_sd) {
  ZZ_CHECK(h);
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  const int64_t CG = (int64_t)h->C * h->G;
  k_init_state<<<gene_grid(h, h->G, h->C), BLOCK, 0, h->stream>>>(h->G, h->cfg.gene_offset, h->cfg.seed, step, rowvar_log_mean, rowvar_sd, h->Xg,
                                                                  h->gl, h->perm, h->nd_mask, h->xbar, h->xss, h->hyper, h->Yg, h->vg, h->ty, h->tv, h->flags);
  ZZ_LAUNCHED(h);
  k_window_init<<<gene_grid(h, CG), BLOCK, 0, h->stream>>>(CG, h->win_y, h->win_v);     // Yg_trace, variance_g_trace: I:460, I:466
  ZZ_LAUNCHED(h);
  h->lps_valid = false; h->stats_fresh = false; h->caches_valid = false;
  return zz_refresh_caches(h);                                        // I:473, I:504
}

int zz_get_alloc_trace(zz_handle *h, int chain, uint32_t *out) {
  ZZ_CHECK(h);
  if (int rc = chain_ok(h, chain)) return rc;
  if (!out) return fail(h, ZZ_ERR_INVALID, "out is NULL");
  const size_t n = (size_t)h->G * (h->K + 1);
  (void)n;
  return down_g(h, out, h->trace + (size_t)chain * n, h->G, h->K + 1);
}

int zz_tune(zz_handle *h, double target) {
  ZZ_CHECK(h);
  const int64_t CG = (int64_t)h->C * h->G;
  k_tune<<<gene_grid(h, CG), BLOCK, 0, h->stream>>>(CG, target, h->win_y, h->win_v, h->ty, h->tv);
  ZZ_LAUNCHED(h);
  return ZZ_OK;
}

int zz_nan_flag(zz_handle *h, int *flag) {
  ZZ_CHECK(h);
  if (!flag) return fail(h, ZZ_ERR_INVALID, "flag is NULL");
  ZZ_CUDA(h, cudaMemcpyAsync(flag, h->nan_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZZ_OK;
}

int zz_sweep_host(zz_handle *h, uint64_t step, int tune, double *Yg, double *variance_g, int32_t *ai, int32_t *within,
                  int32_t *spike, double *XgLikelihood, double *variance_g_probability, const double *ty, const double *tv,
                  int64_t *h2d_bytes, int64_t *d2h_bytes) {
  ZZ_CHECK(h);
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  if (!Yg || !variance_g || !ai || !within || !spike || !XgLikelihood || !variance_g_probability || !ty || !tv)
    return fail(h, ZZ_ERR_INVALID, "NULL state pointer");
  const int64_t G = h->G;
  int rc;
  if (!h->pipe_ready) {
    ZZ_CUDA(h, cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    ZZ_CUDA(h, cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
    ZZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_entry, cudaEventDisableTiming));
    for (int k = 0; k < ZZ_HOST_MAX_CHUNKS; ++k) {
      ZZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_up[k], cudaEventDisableTiming));
      ZZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
    }
    h->pipe_ready = true;
  }
  if (!h->inv_perm) {                                   // caller gene -> device slot
    std::vector<int32_t> inv((size_t)G);
    for (int64_t g = 0; g < G; ++g) inv[(size_t)h->perm_h[(size_t)g]] = (int32_t)g;
    ZZ_CUDA(h, cudaMalloc(&h->inv_perm, sizeof(int32_t) * G));
    ZZ_CUDA(h, cudaMemcpyAsync(h->inv_perm, inv.data(), sizeof(int32_t) * G, cudaMemcpyHostToDevice, h->stream));
    ZZ_CUDA(h, cudaStreamSynchronize(h->stream));      // `inv` is a local
  }
  if ((rc = ensure_fasth(h))) return rc;
  // Three streams: copies in (s_up), kernels (the handle's stream), copies out (s_down); PCIe is full duplex, so in steady state
  // chunk k+1 travels up and chunk k-1 travels down while chunk k is swept.  Staging planes are whole-G arrays in caller order
  // (inbound in `stage`, outbound in `scratch`), each chunk owns its own sub-range: no buffer is reused within a call.
  int nchunks = (int)((G + 131072) / 262144);           // ~256K genes per chunk: below that the per-copy cost of 14 copies per chunk dominates
  if (const char *e = getenv("ZZ_HOST_CHUNKS")) nchunks = atoi(e);
  if (nchunks < 1) nchunks = 1;
  if (nchunks > ZZ_HOST_MAX_CHUNKS) nchunks = ZZ_HOST_MAX_CHUNKS;
  const int64_t per = ((G + nchunks - 1) / nchunks + 31) / 32 * 32;
  HostChunk hc;
  hc.inv = h->inv_perm; hc.ty = h->ty; hc.tv = h->tv;
  double *up8 = h->stage, *dn8 = h->scratch;
  int32_t *up4 = reinterpret_cast<int32_t *>(up8 + 4 * G), *dn4 = reinterpret_cast<int32_t *>(dn8 + 4 * G);
  hc.sYg = up8; hc.svg = up8 + G; hc.sty = up8 + 2 * G; hc.stv = up8 + 3 * G; hc.sai = up4; hc.sw = up4 + G; hc.ss = up4 + 2 * G;
  hc.dYg = dn8; hc.dvg = dn8 + G; hc.dxl = dn8 + 2 * G; hc.dvp = dn8 + 3 * G; hc.dai = dn4; hc.dw = dn4 + G; hc.ds = dn4 + 2 * G;
  {
    // Caller arrays in pinned, device-mapped host memory (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory()): ONE kernel
    // reads them, sweeps, and writes them back over PCIe itself — no staging copies, both directions busy at once.
    void *hp[9] = {Yg, variance_g, (void *)ty, (void *)tv, ai, within, spike, XgLikelihood, variance_g_probability};
    void *dp[9];
    bool mapped = getenv("ZZ_HOST_NO_ZEROCOPY") == nullptr;
    for (int j = 0; j < 9 && mapped; ++j) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, hp[j]) != cudaSuccess) { cudaGetLastError(); mapped = false; break; }
      if (at.type != cudaMemoryTypeHost || !at.devicePointer) mapped = false;
      dp[j] = at.devicePointer;
    }
    if (mapped) {
      hc.c0 = 0; hc.n = G;
      hc.sYg = (const double *)dp[0]; hc.svg = (const double *)dp[1]; hc.sty = (const double *)dp[2]; hc.stv = (const double *)dp[3];
      hc.sai = (const int32_t *)dp[4]; hc.sw = (const int32_t *)dp[5]; hc.ss = (const int32_t *)dp[6];
      hc.dYg = (double *)dp[0]; hc.dvg = (double *)dp[1]; hc.dxl = (double *)dp[7]; hc.dvp = (double *)dp[8];
      hc.dai = (int32_t *)dp[4]; hc.dw = (int32_t *)dp[5]; hc.ds = (int32_t *)dp[6];
      const double *nd[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
      uint8_t *ne[4] = {nullptr, nullptr, nullptr, nullptr};
      h->lps_valid = false;
      if ((rc = launch_sweep_fast(h, step, tune, ZZ_MOVE_FUSED, nd, ne, 0, G, h->inv_perm, &hc))) return rc;
      ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
      h->lps_valid = true;
      if (h2d_bytes) *h2d_bytes = G * (4 * 8 + 3 * 4);
      if (d2h_bytes) *d2h_bytes = G * (4 * 8 + 3 * 4);
      return ZZ_OK;
    }
  }
  ZZ_CUDA(h, cudaEventRecord(h->ev_entry, h->stream));   // earlier work on the handle's stream may still use the staging buffers
  ZZ_CUDA(h, cudaStreamWaitEvent(h->s_up, h->ev_entry, 0));
  const double *no_draws[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t *no_dec[4] = {nullptr, nullptr, nullptr, nullptr};
  h->lps_valid = false;                                  // rebuilt chunk by chunk below; valid again once every chunk is through
  int k = 0;
  for (int64_t c0 = 0; c0 < G; c0 += per, ++k) {
    const int64_t n = (G - c0 < per) ? G - c0 : per;
    const double *h8[4] = {Yg, variance_g, ty, tv};
    const int32_t *h4[3] = {ai, within, spike};
    for (int j = 0; j < 4; ++j) ZZ_CUDA(h, cudaMemcpyAsync(up8 + j * G + c0, h8[j] + c0, sizeof(double) * n, cudaMemcpyHostToDevice, h->s_up));
    for (int j = 0; j < 3; ++j) ZZ_CUDA(h, cudaMemcpyAsync(up4 + j * G + c0, h4[j] + c0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->s_up));
    ZZ_CUDA(h, cudaEventRecord(h->ev_up[k], h->s_up));
    ZZ_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_up[k], 0));
    hc.c0 = c0; hc.n = n;
    FastArgs fa;
    fill_sweep_args(h, fa.s, step, tune, ZZ_MOVE_FUSED, no_draws, no_dec, 0, G);
    fill_fast_args(h, fa);
    ZZ_DISPATCH_R(h, k_host_in, gene_grid(h, n), fa, hc);
    ZZ_LAUNCHED(h);
    if ((rc = launch_sweep_fast(h, step, tune, ZZ_MOVE_FUSED, no_draws, no_dec, c0, c0 + n, h->inv_perm))) return rc;
    k_host_out<<<gene_grid(h, n), BLOCK, 0, h->stream>>>(fa.s, hc);
    ZZ_LAUNCHED(h);
    ZZ_CUDA(h, cudaEventRecord(h->ev_done[k], h->stream));
    ZZ_CUDA(h, cudaStreamWaitEvent(h->s_down, h->ev_done[k], 0));
    double *o8[4] = {Yg, variance_g, XgLikelihood, variance_g_probability};
    int32_t *o4[3] = {ai, within, spike};
    for (int j = 0; j < 4; ++j) ZZ_CUDA(h, cudaMemcpyAsync(o8[j] + c0, dn8 + j * G + c0, sizeof(double) * n, cudaMemcpyDeviceToHost, h->s_down));
    for (int j = 0; j < 3; ++j) ZZ_CUDA(h, cudaMemcpyAsync(o4[j] + c0, dn4 + j * G + c0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->s_down));
  }
  ZZ_CUDA(h, cudaStreamSynchronize(h->s_down));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  h->lps_valid = true;
  if (h2d_bytes) *h2d_bytes = G * (4 * 8 + 3 * 4);
  if (d2h_bytes) *d2h_bytes = G * (4 * 8 + 3 * 4);
  return ZZ_OK;
}

// ---- device-resident generation loop (zz_chain.cuh, zz_persist.cuh) ---------------------------------------------------
static void win_pack(const uint8_t *w, Win100 &o) {          // trace[0] oldest ... trace[99] newest -> bit 0 = newest
  o.lo = 0; o.hi = 0;
  for (int j = 0; j < 100; ++j) if (w[99 - j]) { if (j < 64) o.lo |= 1ULL << j; else o.hi |= 1ULL << (j - 64); }
}
static void win_unpack(const Win100 &w, uint8_t *o) {
  for (int j = 0; j < 100; ++j) o[99 - j] = (uint8_t)((j < 64 ? (w.lo >> j) : (w.hi >> (j - 64))) & 1ULL);
}

int zz_chain_begin(zz_handle *h, const zz_priors *priors, const zz_chain_scalars *scalars) {
  ZZ_CHECK(h);
  if (!priors || !scalars) return fail(h, ZZ_ERR_INVALID, "NULL argument");
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  const int C = h->C;
  for (int c = 1; c < C; ++c)
    if (scalars[c].gen != scalars[0].gen || scalars[c].step != scalars[0].step || scalars[c].n_samples != scalars[0].n_samples)
      return fail(h, ZZ_ERR_INVALID, "zz_chain_begin: batched chains advance in lock-step (same gen, step, n_samples)");
  if (!h->devchain) ZZ_CUDA(h, cudaMalloc(&h->devchain, sizeof(ChainDev) * C));
  std::vector<ChainDev> tmp((size_t)C);
  memset(tmp.data(), 0, sizeof(ChainDev) * C);
  for (int c = 0; c < C; ++c) {
    ChainDev &d = tmp[c]; const zz_chain_scalars &sc = scalars[c];
    d.pr = *priors; d.chain = c;
    memcpy(d.active_means_dif, sc.active_means_dif, sizeof d.active_means_dif);
    d.t_im = sc.t_im; d.t_av = sc.t_av; memcpy(d.t_am, sc.t_am, sizeof d.t_am);
    d.t_s0 = sc.t_s0; d.t_s1 = sc.t_s1; d.t_tau = sc.t_tau; d.t_s0tau = sc.t_s0tau; memcpy(d.t_alpha, sc.t_alpha, sizeof d.t_alpha);
    win_pack(sc.w_im, d.w_im); win_pack(sc.w_av, d.w_av); win_pack(sc.w_s0, d.w_s0); win_pack(sc.w_s1, d.w_s1);
    win_pack(sc.w_tau, d.w_tau); win_pack(sc.w_s0tau, d.w_s0tau);
    for (int k = 0; k < ZZ_MAX_COMP; ++k) win_pack(sc.w_am[k], d.w_am[k]);
    for (int r = 0; r < ZZ_MAX_LIBS; ++r) win_pack(sc.w_alpha[r], d.w_alpha[r]);
    d.hyper_ctr = sc.hyper_ctr; memcpy(d.hyper_buf, sc.hyper_buf, sizeof d.hyper_buf);
    d.step = sc.step; d.gen = sc.gen; d.n_samples = sc.n_samples;
  }
  cudaError_t e = cudaMemcpyAsync(h->devchain, tmp.data(), sizeof(ChainDev) * C, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, ZZ_ERR_CUDA, std::string("zz_chain_begin: ") + cudaGetErrorString(e));
  h->chain_gen = scalars->gen; h->chain_step = scalars->step; h->chain_samples = scalars->n_samples;
  h->chain_ready = true;
  return ZZ_OK;
}

}  // extern "C"
namespace {

// the two CTA shapes of k_generations (see the kernel): 512 threads for shards up to GEN_LARGE_MIN genes and for batched chains, 640 above
// CTA shapes of k_generations (profiles/r2_summary.md): one CTA per SM of 512 threads (128 registers), of 640 threads (96 registers)
// for one chain over a large shard, and three CTAs per SM of 256 threads (80 registers) for whole burn-in generations of batched
// chains, whose scalar stages then overlap other chains' tiles on the same SM (64 chains x 60 k x 4: 264 against 289 us per burn-in
// generation; sampling generations, without the window streams, are 242 us on 512 x 1 against 245 us)
constexpr int GEN_BLK_S = SWEEP_BLOCK, GEN_BLK_L = 640, GEN_BLK_C = 256;
template <int RT, int BLK> const void *generations_fn(int K) {
  switch (K) {
    case 1: return (const void *)k_generations<RT, 1, BLK>;
    case 2: return (const void *)k_generations<RT, 2, BLK>;
    case 3: return (const void *)k_generations<RT, 3, BLK>;
    default: return nullptr;
  }
}
const void *generations_fn(int R, int K, int blk) {
#ifdef ZZ_DEV_R16
  return blk == GEN_BLK_L ? generations_fn<16, GEN_BLK_L>(K) : blk == GEN_BLK_C ? generations_fn<16, GEN_BLK_C>(K) : generations_fn<16, GEN_BLK_S>(K);
#else
#define ZZ_GEN_FN(RT_) (blk == GEN_BLK_L ? generations_fn<RT_, GEN_BLK_L>(K) : blk == GEN_BLK_C ? generations_fn<RT_, GEN_BLK_C>(K) : generations_fn<RT_, GEN_BLK_S>(K))
  switch (R) { case 2: return ZZ_GEN_FN(2); case 4: return ZZ_GEN_FN(4); case 8: return ZZ_GEN_FN(8); case 16: return ZZ_GEN_FN(16); default: return ZZ_GEN_FN(0); }
#undef ZZ_GEN_FN
#endif
}

// whether k_generations serves this handle: straight-line tile (K <= 3, no variance_g upper bound), library loop unrolled
const char *generations_unsupported(zz_handle *h) {
  if (h->K > 3) return "more than 3 active components";
#ifdef ZZ_DEV_R16
  if (h->R != 16) return "ZZ_DEV_R16 build";
#endif
  for (int c = 0; c < h->C; ++c) if (h->hyper_host[c].variance_g_upper_bound != INFINITY) return "finite variance_g_upper_bound (its plnorm term has no closed form in the sufficient statistics)";
  return nullptr;
}

// one cooperative launch: `ngen` generations (hyper_moves) or fused sweeps with frozen hyper-parameters (!hyper_moves)
int launch_generations(zz_handle *h, uint64_t step0, int ngen, int tune, int hyper_moves, int step_stride, int do_tune_all_last, double target) {
  if (ngen <= 0) return ZZ_OK;
  const int C = h->C, nstat = zz_suffstats_len(h->K);
  if (!h->genctl) {
    ZZ_CUDA(h, cudaMalloc(&h->genctl, sizeof(GenCtl) * C));
    ZZ_CUDA(h, cudaMalloc(&h->genacc, sizeof(long long) * (size_t)C * nstat));
    ZZ_CUDA(h, cudaMalloc(&h->dlt, sizeof(double) * (size_t)C * h->G));
    ZZ_CUDA(h, cudaMemsetAsync(h->dlt, 0, sizeof(double) * (size_t)C * h->G, h->stream));
  }
  if (int rc = ensure_lps(h)) return rc;
  static const int64_t large_min = getenv("ZZ_GEN_LARGE_MIN") ? atoll(getenv("ZZ_GEN_LARGE_MIN")) : 700000;
  static const int no_blk_c = getenv("ZZ_GEN_NO_BLK_C") ? 1 : 0;   // measurement switch
  const int blk = (C == 1 && h->G >= large_min) ? GEN_BLK_L : (C > 1 && hyper_moves && tune && !no_blk_c && GEN_BLK_C != GEN_BLK_S) ? GEN_BLK_C : GEN_BLK_S;
  const size_t smem = sizeof(double) * (nstat + 10) * blk;
  const void *fn = generations_fn(h->R, h->K, blk);
  if (!fn) return fail(h, ZZ_ERR_INVALID, "k_generations: more than 3 active components");
  ZZ_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int &gen_blocks = blk == GEN_BLK_L ? h->gen_blocks_l : blk == GEN_BLK_C ? h->gen_blocks_c : h->gen_blocks;
  if (gen_blocks == 0) {    // co-resident CTAs of k_generations (cooperative launch): from the occupancy calculator
    int per_sm = 0, sms = 0;
    ZZ_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, blk, smem));
    ZZ_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    if (per_sm < 1) return fail(h, ZZ_ERR_CUDA, "k_generations does not fit on an SM");
    gen_blocks = per_sm * sms;
    if (const char *e2 = getenv("ZZ_GEN_BLOCKS")) { const int n = atoi(e2); if (n >= 1 && n < gen_blocks) gen_blocks = n; }   // tests: a smaller grid
  }
  const int64_t ntiles = (h->G + 31) / 32, need = (ntiles + blk / 32 - 1) / (blk / 32);
  // one chain: one more CTA, the steward of the scalar stage (zz_persist.cuh); batched chains overlap each other's scalar stages
  // instead, the last CTA of a chain to finish its tiles runs that chain's stage
  static const int no_steward = getenv("ZZ_GEN_NO_STEWARD") ? 1 : 0;   // measurement switch
  const int steward = (C == 1 && !no_steward && gen_blocks >= 2) ? 1 : 0;
  int64_t per_chain = gen_blocks / C;
  if (per_chain < 1) return fail(h, ZZ_ERR_INVALID, "too many chains for the co-resident grid of the device-resident loop");
  (void)need;
  if (per_chain > ntiles + steward) per_chain = ntiles + steward;   // never more worker CTAs than tiles (tiles are dealt CTA-minor)
  GenArgs ga;
  memset(&ga, 0, sizeof ga);
  const double *no_draws[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t *no_dec[4] = {nullptr, nullptr, nullptr, nullptr};
  fill_sweep_args(h, ga.fa.s, step0, tune, ZZ_MOVE_FUSED, no_draws, no_dec, 0, h->G);
  fill_fast_args(h, ga.fa);
  ga.fa.do_spike_move = 1;
  ga.dlt = h->dlt; ga.hyper = h->hyper; ga.chain = h->devchain; ga.fasth_rw = h->fasth; ga.ctl = h->genctl; ga.acc = h->genacc;
  ga.ngen = ngen; ga.hyper_moves = hyper_moves; ga.do_tune_all_last = do_tune_all_last; ga.step_stride = step_stride; ga.target = target;
  ga.stats_out = h->stats_target;
  static const int dyn_tiles = getenv("ZZ_GEN_STATIC_TILES") ? 0 : 1;   // measurement switch (profiles/r2_summary.md)
  ga.dyn_tiles = dyn_tiles;
  ga.steward = steward;
  static const long long timeout_s = getenv("ZZ_GEN_TIMEOUT_S") ? atoll(getenv("ZZ_GEN_TIMEOUT_S")) : 20;
  ga.timeout_ns = timeout_s * 1000000000LL;                          // grid barrier of one chain; the peer exchange: ZZ_COMM_TIMEOUT_S
  if (h->comm_ready && h->comm_world > 1) {
    ga.do_allreduce = 1;
    fill_comm(h, ga.cm, nstat, comm_take_epochs(h, (unsigned long long)ngen));   // generation gi exchanges under epoch ga.cm.epoch + gi + 1
  }
  // diagnostics: ZZ_GEN_TRACE=<file> dumps per-CTA time stamps of the first 16 generations of every launch (this call then waits for the launch)
  static const char *trace_path = getenv("ZZ_GEN_TRACE");
  unsigned long long *trace_dev = nullptr;
  const int trace_gens = ngen < 16 ? ngen : 16;
  const size_t trace_words = (size_t)trace_gens * per_chain * 4;
  if (trace_path && C == 1) {
    ZZ_CUDA(h, cudaMalloc(&trace_dev, trace_words * 8));
    ZZ_CUDA(h, cudaMemsetAsync(trace_dev, 0, trace_words * 8, h->stream));
    ga.trace = trace_dev; ga.trace_gens = trace_gens;
  }
  ZZ_CUDA(h, cudaMemsetAsync(h->genctl, 0, sizeof(GenCtl) * C, h->stream));
  ZZ_CUDA(h, cudaMemsetAsync(h->genacc, 0, sizeof(long long) * (size_t)C * nstat, h->stream));
  const dim3 grid((unsigned)per_chain, (unsigned)C);
  void *kargs[] = {&ga};
  cudaError_t e = cudaLaunchCooperativeKernel(fn, grid, dim3(blk), kargs, smem, h->stream);
  ZZ_CUDA(h, e);
  h->launches++;
  h->caches_valid = false; h->stats_fresh = false;
  if (trace_dev) {
    std::vector<unsigned long long> buf(trace_words);
    ZZ_CUDA(h, cudaMemcpyAsync(buf.data(), trace_dev, trace_words * 8, cudaMemcpyDeviceToHost, h->stream));
    ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(trace_dev);
    if (FILE *f = fopen(trace_path, "ab")) {
      const unsigned long long hdr[4] = {0x7a7a747261636531ull, (unsigned long long)trace_gens, (unsigned long long)per_chain, (unsigned long long)steward};
      fwrite(hdr, 8, 4, f); fwrite(buf.data(), 8, trace_words, f); fclose(f);
    }
  }
  return ZZ_OK;
}

}  // namespace
extern "C" {

int zz_debug_gen_times(zz_handle *h, double *out2) {   // diagnostic: ns the releasing CTA of chain 0 spent in the last launch {ticket -> release, scalar stage, its 6 phases}
  ZZ_CHECK(h);
  if (!h->genctl || !out2) return fail(h, ZZ_ERR_STATE, "no device-resident launch yet");
  GenCtl c;
  ZZ_CUDA(h, cudaMemcpyAsync(&c, h->genctl, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  out2[0] = (double)c.ns_leader; out2[1] = (double)c.ns_stage;
  for (int j = 0; j < 6; ++j) out2[2 + j] = (double)c.ns_phase[j];
  return ZZ_OK;
}

int zz_run_sweeps(zz_handle *h, uint64_t step, int nsweeps, int tune) {
  ZZ_CHECK(h);
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  if (nsweeps < 0) return fail(h, ZZ_ERR_INVALID, "zz_run_sweeps: bad arguments");
  if (const char *why = generations_unsupported(h)) return fail(h, ZZ_ERR_INVALID, std::string("zz_run_sweeps: ") + why);
  if (int rc = ensure_fasth(h)) return rc;
  return launch_generations(h, step, nsweeps, tune, 0, 1, 0, 0.44);
}

int zz_run_generations(zz_handle *h, int ngen, int tune, int sample_frequency, double target) {
  ZZ_CHECK(h);
  if (!h->chain_ready) return fail(h, ZZ_ERR_STATE, "zz_run_generations: call zz_chain_begin first");
  if (ngen < 0 || (!tune && sample_frequency < 1)) return fail(h, ZZ_ERR_INVALID, "zz_run_generations: bad arguments");
  if (h->hyper_host_stale) {                    // the unsupported-model test below reads the host copy
    ZZ_CUDA(h, cudaMemcpyAsync(h->hyper_host.data(), h->hyper, sizeof(zz_hyper) * h->C, cudaMemcpyDeviceToHost, h->stream));
    ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
    h->hyper_host_stale = false;
  }
  if (const char *why = generations_unsupported(h)) return fail(h, ZZ_ERR_INVALID, std::string("zz_run_generations: ") + why + " — use the host-driven schedule");
  const int K = h->K, R = h->R, C = h->C;
  if (C > 1 && h->sink_on) return fail(h, ZZ_ERR_INVALID, "zz_run_generations: the sample sink serves a one-chain handle");
  if (!tune && h->sink_on) {                      // every sample of this call needs a parameter row
    int64_t need = 0;
    for (int i = 0; i < ngen; ++i) need += ((h->chain_gen + i) % sample_frequency == 0);
    if (h->sink_prow + need > h->sink.param_capacity)
      return fail(h, ZZ_ERR_INVALID, "zz_run_generations: the sample sink has too few parameter rows for this call");
  }
  double *saved_target = h->stats_target;
  h->stats_target = nullptr;                      // the statistics of a generation are consumed on the device
  int rc = ZZ_OK;
  if (!tune && h->chain_samples == 0 && ngen > 0) {   // M:142-145: the trace starts at the current allocation
    k_alloc_trace<<<gene_grid(h, h->G, C), BLOCK, 0, h->stream>>>(h->G, K, h->flags, h->trace);
    h->launches++; h->chain_samples = 1;
  }
  // Segments end at the generations after which something other than the next sweep looks at the per-gene state: a sample
  // (allocation trace, log rows: M:188-220) or tune_all (B:146).  Everything in between is ONE cooperative launch.
  int done = 0;
  while (done < ngen && rc == ZZ_OK) {
    int len = 0; bool sampled = false, tune_pt = false;
    while (done + len < ngen) {
      const int64_t g = h->chain_gen + len;
      ++len;
      sampled = !tune && g % sample_frequency == 0;
      tune_pt = tune && (g + 1) % 400 == 0;
      if (sampled || tune_pt || len >= 100000) break;
    }
    k_gen_prepare<<<(C + 63) / 64, 64, 0, h->stream>>>(C, R, tune, 1, h->cfg.seed, h->hyper, h->devchain, h->fasth);
    h->launches++; h->fasth_valid = true;
    if ((rc = launch_generations(h, h->chain_step, len, tune, 1, 10, tune_pt ? 1 : 0, target))) break;
    k_apply_dlt<<<gene_grid(h, h->G, C), BLOCK, 0, h->stream>>>(h->G, h->devchain, h->flags, h->dlt, h->lps);
    h->launches++;
    h->chain_gen += len; h->chain_step += 10ull * (uint64_t)len; done += len;
    if (tune_pt) { const int64_t CG = (int64_t)C * h->G; k_tune<<<gene_grid(h, CG), BLOCK, 0, h->stream>>>(CG, target, h->win_y, h->win_v, h->ty, h->tv); h->launches++; }
    if (sampled) {
      k_alloc_trace<<<gene_grid(h, h->G, C), BLOCK, 0, h->stream>>>(h->G, K, h->flags, h->trace); h->launches++; h->chain_samples++;
      if (h->sink_on) {                           // M:202-220: lnl and the log rows of this sample, written by the device
        const int64_t gen = h->chain_gen - 1;
        const int nblk = red_blocks(h->G);
        RedArgs ra; fill_red(h, ra, 0, RED_LNL);
        ra.partials = h->partials;
        k_reduce<<<dim3(nblk, 1, 1), BLOCK, 0, h->stream>>>(ra);
        k_reduce_final<<<1, BLOCK, 0, h->stream>>>(h->partials, nblk, 2, h->red_out);
        h->launches += 2;
        const bool yv = h->sink_y && ((gen / sample_frequency) % h->sink.yv_every == 0) && h->sink_yrow < h->sink.yv_capacity;
        const int plen = zz_param_row_len(R, K);
        const int64_t nc = h->sink.n_candidates;
        int64_t nb = yv ? (nc + BLOCK - 1) / BLOCK : 1;
        if (nb < 1) nb = 1;
        if (nb > 4 * 148) nb = 4 * 148;
        k_sample_row<<<(unsigned)nb, BLOCK, 0, h->stream>>>((double)gen, h->hyper, h->red_out, R, K, h->sink_p + h->sink_prow * plen,
                                                            h->cand_slots, nc, h->Yg, h->vg, h->flags,
                                                            yv ? h->sink_y + h->sink_yrow * (1 + nc) : nullptr, yv ? h->sink_v + h->sink_yrow * (1 + nc) : nullptr);
        h->launches++; h->sink_prow++; if (yv) h->sink_yrow++;
      }
    }
  }
  h->stats_target = saved_target;
  h->stats_fresh = false;
  if (ngen > 0) { h->hyper_host_stale = true; h->fasth_valid = true; }
  if (rc) return rc;
  ZZ_CUDA(h, cudaGetLastError());
  return ZZ_OK;
}

int zz_chain_read(zz_handle *h, zz_chain_scalars *scalars, zz_hyper *hyper) {
  ZZ_CHECK(h);
  if (!h->chain_ready) return fail(h, ZZ_ERR_STATE, "zz_chain_read: call zz_chain_begin first");
  std::vector<ChainDev> tmp((size_t)h->C);
  ZZ_CUDA(h, cudaMemcpyAsync(h->hyper_host.data(), h->hyper, sizeof(zz_hyper) * h->C, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(tmp.data(), h->devchain, sizeof(ChainDev) * h->C, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));   // the host copy of the hyper-parameters is current again
  h->hyper_host_stale = false;
  for (int c = 0; c < h->C; ++c) {
    if (scalars) {
      zz_chain_scalars &sc = scalars[c]; const ChainDev &d = tmp[c];
      memset(&sc, 0, sizeof sc);
      memcpy(sc.active_means_dif, d.active_means_dif, sizeof sc.active_means_dif);
      sc.t_im = d.t_im; sc.t_av = d.t_av; memcpy(sc.t_am, d.t_am, sizeof sc.t_am);
      sc.t_s0 = d.t_s0; sc.t_s1 = d.t_s1; sc.t_tau = d.t_tau; sc.t_s0tau = d.t_s0tau; memcpy(sc.t_alpha, d.t_alpha, sizeof sc.t_alpha);
      win_unpack(d.w_im, sc.w_im); win_unpack(d.w_av, sc.w_av); win_unpack(d.w_s0, sc.w_s0); win_unpack(d.w_s1, sc.w_s1);
      win_unpack(d.w_tau, sc.w_tau); win_unpack(d.w_s0tau, sc.w_s0tau);
      for (int k = 0; k < ZZ_MAX_COMP; ++k) win_unpack(d.w_am[k], sc.w_am[k]);
      for (int r = 0; r < ZZ_MAX_LIBS; ++r) win_unpack(d.w_alpha[r], sc.w_alpha[r]);
      sc.hyper_ctr = d.hyper_ctr; memcpy(sc.hyper_buf, d.hyper_buf, sizeof sc.hyper_buf);
      sc.step = d.step; sc.gen = d.gen; sc.n_samples = h->chain_samples;
    }
    if (hyper) hyper[c] = h->hyper_host[c];
  }
  return ZZ_OK;
}

// ---- sample-time outputs written by the device (zz_set_sample_sink) -------------------------------------------------
int zz_param_row_len(int R, int K) { return 9 + R + 3 * K; }

void *zz_pinned_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void zz_pinned_free(void *p) { if (p) cudaFreeHost(p); }

int zz_set_sample_sink(zz_handle *h, const zz_sample_sink *sink) {
  ZZ_CHECK(h);
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));        // rows of earlier calls are complete before the rings change
  h->sink_on = false; h->sink_prow = h->sink_yrow = 0;
  if (h->cand_slots) { cudaFree(h->cand_slots); h->cand_slots = nullptr; }
  if (!sink) return ZZ_OK;
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  if (!sink->param_rows || sink->param_capacity < 1) return fail(h, ZZ_ERR_INVALID, "zz_set_sample_sink: no parameter rows");
  const bool yv = sink->yg_rows && sink->vg_rows && sink->yv_capacity > 0 && sink->n_candidates > 0;
  if (yv && (!sink->candidates || sink->yv_every < 1)) return fail(h, ZZ_ERR_INVALID, "zz_set_sample_sink: candidates / yv_every");
  auto dev_ptr = [&](double *host, double **out) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { cudaGetLastError(); return false; }
    *out = (double *)at.devicePointer; return true;
  };
  h->sink_y = h->sink_v = nullptr;
  if (!dev_ptr(sink->param_rows, &h->sink_p) || (yv && (!dev_ptr(sink->yg_rows, &h->sink_y) || !dev_ptr(sink->vg_rows, &h->sink_v))))
    return fail(h, ZZ_ERR_INVALID, "zz_set_sample_sink: row buffers must be pinned, device-mapped host memory (zz_pinned_alloc)");
  if (yv) {
    std::vector<int32_t> inv((size_t)h->G), slots((size_t)sink->n_candidates);
    for (int64_t g = 0; g < h->G; ++g) inv[(size_t)h->perm_h[(size_t)g]] = (int32_t)g;
    for (int64_t c = 0; c < sink->n_candidates; ++c) {
      if (sink->candidates[c] < 0 || sink->candidates[c] >= h->G) return fail(h, ZZ_ERR_INVALID, "zz_set_sample_sink: candidate index out of range");
      slots[(size_t)c] = inv[(size_t)sink->candidates[c]];
    }
    ZZ_CUDA(h, cudaMalloc(&h->cand_slots, sizeof(int32_t) * sink->n_candidates));
    ZZ_CUDA(h, cudaMemcpy(h->cand_slots, slots.data(), sizeof(int32_t) * sink->n_candidates, cudaMemcpyHostToDevice));
  }
  h->sink = *sink;
  if (!yv) { h->sink.n_candidates = 0; h->sink.yv_capacity = 0; }
  h->sink_on = true;
  return ZZ_OK;
}

int zz_sample_rows_written(zz_handle *h, int64_t *param_rows, int64_t *yv_rows) {
  ZZ_CHECK(h);
  if (param_rows) *param_rows = h->sink_prow;
  if (yv_rows) *yv_rows = h->sink_yrow;
  return ZZ_OK;
}

// ---- posterior-predictive discrepancies (k_post_pred) ---------------------------------------------------------------
int zz_post_predictive(zz_handle *h, uint64_t step, const double *bins, int nbins, double *out) {
  return zz_post_predictive_sharded(h, step, bins, nbins, nullptr, nullptr, out);
}

// The statistics are functionals of histograms over ALL genes: with gene shards the per-shard histograms (integer counts, fixed-point
// model weights) and the per-library sums are summed over the shards by the caller's `allreduce` (host buffers, in place) — once after
// pass 1 (the library means that centre pass 2 are global) and once after pass 2.  Every rank then holds the same statistics.
int zz_post_predictive_sharded(zz_handle *h, uint64_t step, const double *bins, int nbins, zz_allreduce_fn allreduce, void *ctx, double *out) {
  ZZ_CHECK(h);
  if (!h->have_data) return fail(h, ZZ_ERR_STATE, "zz_upload_data has not been called");
  if (!bins || !out || nbins < 3 || nbins > 1400) return fail(h, ZZ_ERR_INVALID, "zz_post_predictive: need 3 <= nbins <= 1400 (shared-memory histograms)");
  const int R = h->R, K = h->K, nb = nbins, Hn = nb + 1;
  const int64_t G = h->G;
  const size_t n_u32 = (size_t)4 * R * Hn + 2 * Hn, n_u64 = (size_t)R * Hn;
  const size_t bytes = n_u64 * 8 + (size_t)nb * 8 + n_u32 * 4;
  if (!h->pp_buf || h->pp_nb < nb) {
    if (h->pp_buf) cudaFree(h->pp_buf);
    h->pp_buf = nullptr;
    ZZ_CUDA(h, cudaMalloc(&h->pp_buf, bytes));
    h->pp_nb = nb;
  }
  unsigned long long *md = reinterpret_cast<unsigned long long *>(h->pp_buf);
  double *dbins = reinterpret_cast<double *>(md + n_u64);
  unsigned *u32 = reinterpret_cast<unsigned *>(dbins + nb);
  ZZ_CUDA(h, cudaMemsetAsync(h->pp_buf, 0, bytes, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(dbins, bins, sizeof(double) * nb, cudaMemcpyHostToDevice, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(&h->hyper_host[0], h->hyper, sizeof(zz_hyper), cudaMemcpyDeviceToHost, h->stream));   // may have moved on the device
  PPArgs a;
  memset(&a, 0, sizeof a);
  a.G = G; a.R = R; a.nb = nb; a.pass = 1; a.seed = h->cfg.seed; a.step = step; a.gene_offset = h->cfg.gene_offset;
  a.Xg = h->Xg; a.gl = h->gl; a.Yg = h->Yg; a.vg = h->vg; a.flags = h->flags; a.perm = h->perm; a.hyper = h->hyper; a.bins = dbins;
  a.cd = u32; a.cs = u32 + (size_t)R * Hn; a.cds = u32 + (size_t)2 * R * Hn; a.css = u32 + (size_t)3 * R * Hn;
  a.hy = u32 + (size_t)4 * R * Hn; a.hsy = a.hy + Hn; a.md = md; a.partials = h->partials;
  const int nblk = red_blocks(G);
  const size_t pp_smem = sizeof(double) * nb + (size_t)Hn * (8 + 4 * 4);
  k_post_pred<<<dim3(nblk, R), BLOCK, pp_smem, h->stream>>>(a);
  ZZ_LAUNCHED(h);
  k_reduce_final<<<R, BLOCK, 0, h->stream>>>(h->partials, nblk, 4, h->red_out);
  ZZ_LAUNCHED(h);
  std::vector<double> sums((size_t)4 * R);
  std::vector<unsigned> hu(n_u32);
  std::vector<unsigned long long> hm(n_u64);
  ZZ_CUDA(h, cudaMemcpyAsync(sums.data(), h->red_out, sizeof(double) * 4 * R, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(hu.data(), u32, sizeof(unsigned) * ((size_t)2 * R * Hn), cudaMemcpyDeviceToHost, h->stream));   // cd, cs
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  double G_all = (double)G;
  if (allreduce) {
    sums.push_back(G_all);
    allreduce(ctx, sums.data(), (int64_t)sums.size(), ZZ_DTYPE_F64);
    G_all = sums.back(); sums.pop_back();
    allreduce(ctx, hu.data(), (int64_t)2 * R * Hn, ZZ_DTYPE_U32);
  }
  // library means of the detected values (PostPred:52-57) -> pass 2 bins the mean-centred copies
  std::vector<double> ndd(R), nds(R);
  double gmd = 0, gms = 0, lmd[ZZ_MAX_LIBS], lms[ZZ_MAX_LIBS];
  for (int r = 0; r < R; ++r) {
    double cdn = 0, csn = 0;
    for (int i = 0; i < Hn; ++i) { cdn += hu[(size_t)r * Hn + i]; csn += hu[(size_t)(R + r) * Hn + i]; }
    ndd[r] = cdn; nds[r] = csn;
    lmd[r] = sums[4 * r + 0] / cdn; lms[r] = sums[4 * r + 1] / csn; gmd += lmd[r] / R; gms += lms[r] / R;
  }
  for (int r = 0; r < R; ++r) { a.shift_d[r] = -lmd[r] + gmd; a.shift_s[r] = -lms[r] + gms; }
  a.pass = 2;
  k_post_pred<<<dim3(nblk, R), BLOCK, pp_smem, h->stream>>>(a);
  ZZ_LAUNCHED(h);
  ZZ_CUDA(h, cudaMemcpyAsync(hu.data(), u32, sizeof(unsigned) * n_u32, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaMemcpyAsync(hm.data(), md, sizeof(unsigned long long) * n_u64, cudaMemcpyDeviceToHost, h->stream));
  ZZ_CUDA(h, cudaStreamSynchronize(h->stream));
  if (allreduce) {
    allreduce(ctx, hu.data(), (int64_t)n_u32, ZZ_DTYPE_U32);
    allreduce(ctx, hm.data(), (int64_t)n_u64, ZZ_DTYPE_U64);
  }
  // ---- the functionals of the histograms (small: R x nbins), as the reference forms them
  const zz_hyper &hy = h->hyper_host[0];
  const unsigned *cd = hu.data(), *cs = cd + (size_t)R * Hn, *cds = cs + (size_t)R * Hn, *css = cds + (size_t)R * Hn, *hyh = css + (size_t)R * Hn, *hsy = hyh + Hn;
  double n_out = 0;
  for (int i = 0; i < Hn; ++i) n_out += hyh[i];
  auto trapz = [&](const double *x, const double *y, int n) { long double s = 0; for (int i = 1; i < n; ++i) s += (long double)(x[i] - x[i - 1]) * (y[i] + y[i - 1]); return (double)(0.5L * s); };
  auto lower = [&](const unsigned *cnt, const std::vector<double> &ndet, int lo, int hi2) {   // get_lowerLevel_wasserstein, PostPred:120-144
    std::vector<double> mx(nb, 0.0), mm(nb, 0.0), mc(nb), diff(nb);
    for (int r = lo; r < hi2; ++r) {
      long double cm = 0, cx = 0; double maxm = -INFINITY;
      for (int b = 0; b < nb; ++b) { cm += (long double)hm[(size_t)r * Hn + b] / 1099511627776.0L; mc[b] = (double)cm / n_out; if (mc[b] > maxm) maxm = mc[b]; }
      for (int b = 0; b < nb; ++b) { cx += cnt[(size_t)r * Hn + b]; mx[b] += ((double)cx / ndet[r]) / (hi2 - lo); mm[b] += (mc[b] / maxm) / (hi2 - lo); }
    }
    for (int b = 0; b < nb; ++b) diff[b] = fabs(mx[b] - mm[b]);
    return trapz(bins, diff.data(), nb);
  };
  out[0] = lower(cd, ndd, 0, R) - lower(cs, nds, 0, R);                                            // W_L, PostPred:43-44
  for (int r = 0; r < R; ++r) out[3 + r] = lower(cds, ndd, r, r + 1) - lower(css, nds, r, r + 1);  // W_L_r, PostPred:61-66
  {   // get_rumsfeld, PostPred:107-118
    const double c = hy.spike_probability * (1 - hy.weight_active);
    double md_ = 0, ms_ = 0;
    for (int r = 0; r < R; ++r) {
      const double model = (1 - c) * (sums[4 * r + 2] / n_out) + c;
      const double dd = fabs((G_all - ndd[r]) / G_all - model), ds = fabs((G_all - nds[r]) / G_all - model);
      md_ += dd / R; ms_ += ds / R; out[3 + R + r] = dd - ds;
    }
    out[2] = md_ - ms_;
  }
  {   // get_upperLevel_wasserstein (PostPred:146-154) with get_model_cumulative (U:20-40)
    std::vector<double> dens(nb), dy(nb, 0.0), dsy(nb, 0.0);
    const double wi = (1 - hy.weight_active) * (1 - hy.spike_probability), wa = hy.weight_active, s2p = sqrt(2 * 3.14159265358979323846);
    auto dnorm = [&](double x, double mu, double sd) { const double z = (x - mu) / sd; return exp(-0.5 * z * z) / (sd * s2p); };
    for (int b = 0; b < nb; ++b) {
      long double acc = 0;
      for (int k = 0; k < K; ++k) acc += hy.weight_within_active[k] * dnorm(bins[b], hy.active_means[k], sqrt(hy.active_variances[k]));
      dens[b] = (wi * dnorm(bins[b], hy.inactive_means, sqrt(hy.inactive_variances)) + wa * (double)acc) / (wi + wa);
    }
    long double cum = 0, cy = hyh[0], csy = hsy[0];
    for (int b = 1; b < nb; ++b) {
      cum += (long double)(bins[b] - bins[b - 1]) * (dens[b] + dens[b - 1]);
      const double mc = (double)(0.5L * cum);
      cy += hyh[b]; csy += hsy[b];
      dy[b] = fabs((double)cy / n_out - mc); dsy[b] = fabs((double)csy / n_out - mc);
    }
    out[1] = trapz(bins + 1, dy.data() + 1, nb - 1) - trapz(bins + 1, dsy.data() + 1, nb - 1);
  }
  return ZZ_OK;
}

}  // extern "C"
