// zz_persist.cuh — k_generations: whole segments of burnin() / mcmc() generations in ONE cooperative launch (SURVEY §8 f1;
// R/zigzagBurnin.R:64-148, R/zigzagMCMC.R:169-229).  Included by zz_gpu.cu after the sweep kernels (uses their argument structs).
//
// Per generation, every persistent CTA (one per SM, 512 or 640 threads, co-resident by cooperative launch):
//   1. waits until the hyper-parameters of this generation are published, copies the derived constants (FastH) to shared memory;
//   2. sweeps its 32-gene tiles: mhYg -> mhVariance_g -> allocation Gibbs -> mhSpikeAllocation as one straight-line block per
//      tile (same arithmetic as k_sweep_fast's production tile), plus the gene's term of the pending mhP_x likelihood ratio
//      (Pr:209-224) at its final Yg — the separate pass over all genes that move used to need is gone;
//   3. adds its genes' sufficient statistics to the chain's accumulators.  Statistics are FIXED-POINT integers (Q31.32): integer
//      addition is associative, so the totals do not depend on which warp took which tile, on the grid size or on the number of
//      GPUs — results are bit-identical for any sharding, and tiles may be scheduled dynamically;
//   4. takes a ticket.  The LAST CTA to arrive reads the totals, exchanges them with the other GPUs over NVLink peer memory
//      (gene shards: same low-latency inbox format as allreduce_exchange, integer payload), runs every hyper-parameter move of the
//      generation on four of its warps (zz_chain.cuh), derives the next generation's constants and releases the grid.
// Critical path per generation: sweep + one grid barrier (2 us) + the scalar stage; no launch, no host.
#pragma once

namespace {

constexpr double FX_SCALE = 4294967296.0;          // statistics carry 32 fraction bits: per-gene rounding 1.2e-10 absolute
constexpr long long FX_ONE = 1LL << 32;
__device__ __forceinline__ long long to_fx(double x) { return __double2ll_rn(x * FX_SCALE); }

// slots 7 and 8 of the zz_suffstats layout (sum XgLikelihood, sum variance_g_probability: the two caches this kernel does not
// maintain) carry the mhP_x gene sum and its count of non-finite terms instead
constexpr int FXS_PXDELTA = 7, FXS_PXBAD = 8;

// Tiles are handed out dynamically after two static rounds: a 32-gene tile takes a warp ~7 us and the tiles of one sweep differ
// in cost (the detection loop is skipped where p_x saturates; whether it does depends on the proposed Yg), so with a static
// assignment the last warp finishes ~20 us after the first — time every CTA would spend at the generation's grid barrier.
// One shared counter cannot serve 31 250 tickets per 50 us sweep (1.2 ns per same-address atomic, measured: profiles/r2_ubench.txt),
// so the dynamic tiles are dealt round-robin into GEN_NQ queues with a counter each; a warp drains its home queue, then the others.
constexpr int GEN_NQ = 16;
constexpr int GEN_NREL = 16;
struct GenCtl {                     // one per chain, device memory
  unsigned arrive;                  // tickets taken (monotone across launches)
  unsigned release;                 // generations published (monotone); GEN_ABORT = give up (time-out)
  unsigned long long ns_leader, ns_stage;   // diagnostics (zz_debug_gen_times): time the last CTA spent between its ticket and the release / in the scalar stage
  unsigned long long ns_phase[6];           // ... and in its phases: load, draws, chains of moves, install + tune_all, derive + propose, write-back
  unsigned pad_[14];
  unsigned rel[GEN_NREL * 32];      // copies of `release`, one 128-byte line each: CTA b polls rel[b % GEN_NREL], so the 444 pollers of a chain
                                    // spread over 16 L2 lines (measured: no effect on the stage's duration; kept, it costs 16 stores)
  unsigned tile_q[GEN_NQ * 32];     // dynamic tile scheduling: tickets handed out by each tile queue in the current generation, one
                                    // 128-byte line per counter (same-line atomics serialise in one L2 slice)
};
constexpr unsigned GEN_ABORT = 0xFFFFFFFFu;

struct GenArgs {
  FastArgs fa;                      // arrays, tables, fasth; fa.s.step = step of the first generation's sweep
  double *dlt;                      // [C][G] per-gene term of the last mhP_x ratio (added to lps when that move was accepted)
  zz_hyper *hyper; ChainDev *chain; // [C]
  FastH *fasth_rw;                  // == fa.fasth: the leader writes the next generation's constants
  GenCtl *ctl; long long *acc;      // [C] (zeroed by the host before every launch), [C][nstat]
  int ngen, hyper_moves, do_tune_all_last, step_stride; double target;
  double *stats_out;                // [C][nstat] or NULL: the (all-reduced) statistics of the last generation run, as doubles
  int do_allreduce; CommArgs cm;    // cm.epoch = epoch of the first generation - 1
  long long timeout_ns;
  unsigned long long *trace; int trace_gens;   // diagnostics (ZZ_GEN_TRACE): [gen][CTA][4] globaltimer stamps {release seen, tiles start, tiles done, ticket}
  int steward;                      // 1: the last CTA of the (single) chain is the steward of the scalar stage (see k_generations)
  int dyn_tiles;                    // 0: static tile assignment (warp w: tiles w, w + W, ...); 1: static for two rounds, then from the queues
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// integer all-reduce of `n` 64-bit values over the peer inboxes (see allreduce_exchange): every value travels as two words
// {half | epoch << 32}; the sum is exact, so the order of the ranks does not matter.  One thread per (peer, value) polls — a thread
// walking over the peers one after the other paid one L2 round trip per peer AFTER the words had arrived (6 us at 8 GPUs).
// `xch`: shared memory, world * n values.  Returns false on time-out.
__device__ __forceinline__ bool allreduce_exchange_i64(long long *vec /* shared memory, in/out */, long long *xch, int n, const CommArgs &cm, long long timeout_ns) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  const int slot = (int)(cm.epoch & 1ull), nw = 2 * n;
  const unsigned ep = (unsigned)cm.epoch;
  for (int t = threadIdx.x; t < nw * cm.world; t += blockDim.x) {       // one store per thread: (peer, word)
    const int r = t / nw, w = t - r * nw;
    if (r == cm.rank) continue;
    const unsigned long long v = (unsigned long long)vec[w >> 1];
    const unsigned half = (w & 1) ? (unsigned)(v >> 32) : (unsigned)v;
    const unsigned long long word = ((unsigned long long)ep << 32) | half;
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(cm.inbox[r]) + cm.base + ((int64_t)slot * cm.world + cm.rank) * cm.stride + w;
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
  }
  const unsigned long long *in = reinterpret_cast<const unsigned long long *>(cm.inbox[cm.rank]) + cm.base + (int64_t)slot * cm.world * cm.stride;
  const unsigned long long t0 = globaltimer_ns();
  for (int t = threadIdx.x; t < n * cm.world; t += blockDim.x) {
    const int r = t / n, idx = t - r * n;
    long long v = vec[idx];
    if (r != cm.rank) {
      const volatile unsigned long long *p = in + (int64_t)r * cm.stride + 2 * idx;
      unsigned long long a = p[0], b = p[1];
      while ((unsigned)(a >> 32) != ep || (unsigned)(b >> 32) != ep) {
        if ((long long)(globaltimer_ns() - t0) > timeout_ns) { timed_out = 1; break; }
        a = p[0]; b = p[1];
      }
      v = (long long)(((b & 0xFFFFFFFFull) << 32) | (a & 0xFFFFFFFFull));
    }
    xch[t] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    long long sum = 0;
    for (int r = 0; r < cm.world; ++r) sum += xch[r * n + threadIdx.x];
    vec[threadIdx.x] = sum;
  }
  __syncthreads();
  return timed_out == 0;
}


// ---- the scalar half of a generation (zz_chain.cuh) as the persistent kernel runs it ------------------------------------------
// Scratch layout (shared memory; in the last-arriver mode it aliases the first statistic columns of the CTA, which are zero then)
struct StageScr {
  long long *ivec; double *st, *gam, *vg_s, *px_lu; L2Params *l2; zz_hyper *h; ChainDev *c; MoveDraws *md; FastH *F; GammaPre *gp; PreVarg *pv; PreL2 *pl; PrePx *pp; long long *xch;
};
constexpr int MAX_NSTAT_P = 3 * (ZZ_MAX_COMP + 1) + 10;
constexpr int SCR_MD = (2560 + (int)sizeof(ChainDev) + 127) / 128 * 128;
constexpr int SCR_FH = (SCR_MD + MD_MAX * (int)sizeof(MoveDraws) + 127) / 128 * 128;
constexpr int SCR_GP = (SCR_FH + (int)sizeof(FastH) + 127) / 128 * 128;
constexpr int SCR_PV = (SCR_GP + (ZZ_MAX_COMP + 3) * (int)sizeof(GammaPre) + 127) / 128 * 128;
constexpr int SCR_PL = (SCR_PV + (int)sizeof(PreVarg) + 127) / 128 * 128;
constexpr int SCR_PP = (SCR_PL + (int)sizeof(PreL2) + 127) / 128 * 128;
constexpr int SCR_XCH = (SCR_PP + (int)sizeof(PrePx) + 127) / 128 * 128;     // [8 ranks][nstat] parts of the all-reduce
constexpr int SCR_BYTES = SCR_XCH + 8 * MAX_NSTAT_P * 8;
static_assert(MAX_NSTAT_P * 8 <= 512 && sizeof(L2Params) <= 256 && sizeof(zz_hyper) <= 1024, "stage scratch layout");
static_assert(SCR_BYTES <= 8 * 256 * 6, "the scratch must fit in the first statistic columns (nstat >= 13 of them, CTAs of >= 256 threads)");
static_assert(sizeof(ChainDev) % 8 == 0 && sizeof(zz_hyper) % 8 == 0 && sizeof(FastH) % 8 == 0, "copied as 64-bit words");
__device__ __forceinline__ StageScr stage_scr(char *scr) {
  StageScr S;
  S.ivec = reinterpret_cast<long long *>(scr);                 // [nstat] totals, fixed point
  S.st = reinterpret_cast<double *>(scr + 512);                // [nstat] the totals as doubles
  S.gam = reinterpret_cast<double *>(scr + 1024);              // [K + 3] gamma draws
  S.vg_s = reinterpret_cast<double *>(scr + 1152);             // (s0, s1, tau) after the variance-prior moves
  S.px_lu = reinterpret_cast<double *>(scr + 1184);            // log of mhP_x's accept uniform
  S.l2 = reinterpret_cast<L2Params *>(scr + 1280);
  S.h = reinterpret_cast<zz_hyper *>(scr + 1536);
  S.c = reinterpret_cast<ChainDev *>(scr + 2560);
  S.md = reinterpret_cast<MoveDraws *>(scr + SCR_MD);          // [MD_MAX] the prepared draws of the Metropolis moves
  S.F = reinterpret_cast<FastH *>(scr + SCR_FH);               // the next generation's constants, staged (no global read-backs)
  S.gp = reinterpret_cast<GammaPre *>(scr + SCR_GP);           // [K + 3] the statistics-independent part of the gamma draws
  S.pv = reinterpret_cast<PreVarg *>(scr + SCR_PV);            // the statistics-independent parts of the Metropolis moves
  S.pl = reinterpret_cast<PreL2 *>(scr + SCR_PL);
  S.pp = reinterpret_cast<PrePx *>(scr + SCR_PP);              // the next generation's mhP_x proposal
  S.xch = reinterpret_cast<long long *>(scr + SCR_XCH);
  return S;
}
// totals of the generation -> S.st (all-reduced over the gene shards); whole CTA.  false: a peer timed out (fail-stop)
__device__ __forceinline__ bool stage_totals(const StageScr &S, const GenArgs &ga, GenCtl *ctl, long long *gacc, int nstat, int chain, int gi) {
  if ((int)threadIdx.x < nstat) { S.ivec[threadIdx.x] = (long long)__ldcg(reinterpret_cast<const unsigned long long *>(gacc + threadIdx.x)); gacc[threadIdx.x] = 0; }
  if (threadIdx.x < GEN_NQ) ctl->tile_q[threadIdx.x * 32] = 0;
  __syncthreads();
  bool ok = true;
  if (ga.do_allreduce) {
    CommArgs cm = ga.cm;
    cm.epoch += (unsigned long long)gi + 1ull;
    cm.base += (int64_t)chain * 2 * cm.world * cm.stride;
    ok = allreduce_exchange_i64(S.ivec, S.xch, nstat, cm, cm.timeout_ns);
  }
  if ((int)threadIdx.x < nstat) {
    S.st[threadIdx.x] = ok ? (double)S.ivec[threadIdx.x] * (1.0 / FX_SCALE) : NAN;
    if (ga.stats_out) ga.stats_out[(int64_t)chain * nstat + threadIdx.x] = S.st[threadIdx.x];
  }
  if (!ok) {                                            // a peer never showed up: fail-stop (flag bit 2), nobody consumes the stale words
    if (threadIdx.x == 0) { atomicOr(ga.fa.s.nan_flag, 4); __threadfence(); }
    __syncthreads();
    if (threadIdx.x < GEN_NREL) st_release_u32(&ctl->rel[threadIdx.x * 32], GEN_ABORT);
    if (threadIdx.x == 0) st_release_u32(&ctl->release, GEN_ABORT);
  }
  __syncthreads();
  return ok;
}
__device__ __forceinline__ void stage_load(const StageScr &S, const GenArgs &ga, int chain) {      // written by another SM / kernel: through L2
  const uint64_t *s0 = reinterpret_cast<const uint64_t *>(ga.chain + chain), *s1 = reinterpret_cast<const uint64_t *>(ga.hyper + chain);
  uint64_t *d0 = reinterpret_cast<uint64_t *>(S.c), *d1 = reinterpret_cast<uint64_t *>(S.h);
  for (int j = threadIdx.x; j < (int)(sizeof(ChainDev) / 8); j += blockDim.x) d0[j] = __ldcg(s0 + j);
  for (int j = threadIdx.x; j < (int)(sizeof(zz_hyper) / 8); j += blockDim.x) d1[j] = __ldcg(s1 + j);
}
__device__ __forceinline__ void stage_store(const StageScr &S, const GenArgs &ga, int chain) {
  uint64_t *d0 = reinterpret_cast<uint64_t *>(ga.chain + chain), *d1 = reinterpret_cast<uint64_t *>(ga.hyper + chain);
  const uint64_t *s0 = reinterpret_cast<const uint64_t *>(S.c), *s1 = reinterpret_cast<const uint64_t *>(S.h);
  for (int j = threadIdx.x; j < (int)(sizeof(ChainDev) / 8); j += blockDim.x) d0[j] = s0[j];
  for (int j = threadIdx.x; j < (int)(sizeof(zz_hyper) / 8); j += blockDim.x) d1[j] = s1[j];
}
// everything of the generation's scalar stage that needs no statistics: every random draw, one lane per draw
template <int KT>
__device__ __forceinline__ void stage_pre(const StageScr &S, const MathTabs &T, const SweepArgs &a, uint64_t step, uint64_t next_step) {
  const int K = a.K, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const SMath sm{&T};
  if (wid == 0) stage_gibbs_predraw(sm, *S.c, a.seed, step, K, lane, S.gp);
  else if (wid == 4 && lane < MD_AMD + K) make_move_draws(sm, S.md[lane], a.seed, (uint32_t)S.c->chain, step, lane);
  else if (wid == 3 && lane == 0) *S.px_lu = px_log_u(*S.c, a.seed, step);
  else if (wid == 5 && lane == 0) px_propose_pre(*S.c, *S.h, a.seed, next_step, a.R, *S.pp);
  __syncthreads();
  // proposals, Hastings and prior ratios of the Metropolis moves for every outcome of the moves before them
  if (wid == 1 && lane == 0) stage_l2_pre<KT>(sm, *S.c, *S.h, S.md, *S.pl);
  else if (wid == 2 && lane == 0) stage_varg_pre(sm, *S.c, *S.h, S.md, *S.pv);
}
// the moves themselves, from the totals in S.st: new hyper-parameters in S.h, scalar state in S.c, the next sweep's constants in
// S.F and in global memory.  Whole CTA; the caller fences and releases the grid afterwards.
template <int KT>
__device__ __forceinline__ void stage_post(const StageScr &S, const MathTabs &T, const GenArgs &ga, GenCtl *ctl, uint64_t step, bool last, int chain) {
  const SweepArgs &a = ga.fa.s;
  const int K = a.K, R = a.R, K1 = K + 1, tune = a.tune, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const SMath sm{&T};
  ChainDev &c_s = *S.c; zz_hyper &h_s = *S.h; FastH &F_s = *S.F;
  const double *st = S.st;
  unsigned long long tp0 = globaltimer_ns(), tp1;
#ifdef ZZ_GEN_PROFILE   /* per-phase times of the stage (each probe costs a global read-modify-write, ~0.5 us) */
#define ZZ_PHASE(j) do { if (threadIdx.x == 0) { tp1 = globaltimer_ns(); ctl->ns_phase[j] += tp1 - tp0; tp0 = tp1; } } while (0)
#else
#define ZZ_PHASE(j) do { (void)tp0; (void)tp1; (void)ctl; } while (0)
#endif
  // phase A: the Gibbs draws (gamma variates at the counts just reduced)
  if (wid == 0) {
    stage_gibbs_draws(sm, c_s, h_s, a.seed, step, st, K, lane, S.gp, S.gam);
    __syncwarp();
    if (lane == 0) stage_gibbs_combine(h_s, K, S.gam);
  }
  // phase B (does not read what phase A writes): three chains of moves side by side, each on parameters of its own, each installing
  // its results at its end — mhInactiveMeans -> mhActiveMeansDif -> mhSharedVariance (at the means the first two left) | mhTau ->
  // mhSg -> mhS0Tau | the decision of the pending mhP_x and the proposal of the next one
  else if (wid == 1 && lane == 0) {
    double am[KT];
    const double im = stage_im_post(*S.pl, c_s, h_s, S.md, st, K, tune);
    stage_amd_post<KT>(*S.pl, c_s, h_s, S.md, st, tune, am);
    const double av = stage_sv_post<KT>(sm, *S.pl, c_s, h_s, S.md, st, tune, im, am);
    h_s.inactive_means = im;
#pragma unroll
    for (int k = 0; k < KT; ++k) h_s.active_means[k] = am[k];
    if (!isnan(av)) { h_s.inactive_variances = av; for (int k = 0; k < K; ++k) h_s.active_variances[k] = av; }
  }
  else if (wid == 2 && lane == 0) {
    double s0, s1, tau;
    stage_varg_post(*S.pv, c_s, h_s, S.md, st, K, tune, s0, s1, tau);
    h_s.s0 = s0; h_s.s1 = s1; h_s.tau = tau;
  }
  else if (wid == 3 && lane == 0) {
    const int acc = px_decide(c_s, h_s, *S.px_lu, st[3 * K1 + FXS_PXDELTA], st[3 * K1 + FXS_PXBAD], tune);
    if (acc) h_s.alpha_r[c_s.px_lib] = c_s.px_alpha;               // Pr:234-239
    c_s.px_prev_acc = acc; c_s.px_valid = 0;
    if (!last) px_propose_post(*S.pp, c_s, acc, tune);             // the next generation's mhP_x (after a segment: k_gen_prepare)
    c_s.step += (uint64_t)ga.step_stride; c_s.gen += 1;
  }
  __syncthreads();
  ZZ_PHASE(2);
  if (last && ga.do_tune_all_last)                                 // B:146 tune_all, scalar part (the per-gene part follows the launch)
    for (int l = threadIdx.x; l < R + K + 6; l += blockDim.x) tune_all_scalars(c_s, ga.target, K, R, l);
  ZZ_PHASE(3);
  // the next generation's constants, one group of fields per thread
  if (threadIdx.x == DERIVE_FIN_THREAD) { derive_fin(F_s, &h_s); set_px(F_s, h_s, c_s); }
  else derive_fast_par(F_s, &h_s, threadIdx.x);
  __syncthreads();
  {
    uint64_t *dF = reinterpret_cast<uint64_t *>(ga.fasth_rw + chain);
    const uint64_t *sF = reinterpret_cast<const uint64_t *>(&F_s);
    for (int j = threadIdx.x; j < (int)(sizeof(FastH) / 8); j += blockDim.x) dF[j] = sF[j];
  }
  ZZ_PHASE(4);
#undef ZZ_PHASE
}

// The steward CTA of a one-chain launch (see k_generations).  Critical path per generation: last ticket -> totals -> stage_post ->
// release; the state load, every draw and the write-back of the state are off it.
template <int KT>
__device__ __noinline__ void steward_loop(const GenArgs &ga, const MathTabs &T, char *scr, GenCtl *ctl, long long *gacc, int nstat, int chain, unsigned n_workers) {
  __shared__ unsigned sh_ok;
  const SweepArgs &a = ga.fa.s;
  const StageScr S = stage_scr(scr);
  if (ga.hyper_moves) stage_load(S, ga, chain);
  __syncthreads();
  for (int gi = 0; gi < ga.ngen; ++gi) {
    const uint64_t step = a.step + (uint64_t)ga.step_stride * (uint64_t)gi;
    if (ga.hyper_moves) stage_pre<KT>(S, T, a, step, step + (uint64_t)ga.step_stride);
    if (threadIdx.x == 0) {                                          // all tickets of this generation
      const unsigned want = ((unsigned)gi + 1u) * n_workers;
      const unsigned long long t0 = globaltimer_ns();
      unsigned ok = 1;
      while (ld_acquire_u32(&ctl->arrive) < want)
        if ((long long)(globaltimer_ns() - t0) > ga.timeout_ns) { ok = 0; break; }
      sh_ok = ok;
    }
    __syncthreads();
    if (!sh_ok) {
      if (threadIdx.x == 0) { atomicOr(a.nan_flag, 8); __threadfence(); }
      __syncthreads();
      if (threadIdx.x < GEN_NREL) st_release_u32(&ctl->rel[threadIdx.x * 32], GEN_ABORT);
      if (threadIdx.x == 0) st_release_u32(&ctl->release, GEN_ABORT);
      break;
    }
    const unsigned long long t_lead0 = globaltimer_ns();
    if (!stage_totals(S, ga, ctl, gacc, nstat, chain, gi)) break;
    const unsigned long long t_stage0 = globaltimer_ns();
    if (ga.hyper_moves) stage_post<KT>(S, T, ga, ctl, step, gi == ga.ngen - 1, chain);
    __threadfence();
    __syncthreads();
    if (threadIdx.x < GEN_NREL) st_release_u32(&ctl->rel[threadIdx.x * 32], (unsigned)gi + 1u);
    if (threadIdx.x == 0) {
      const unsigned long long t1 = globaltimer_ns();
      if (ga.trace && gi < ga.trace_gens) { unsigned long long *tr = ga.trace + ((size_t)gi * gridDim.x + blockIdx.x) * 4; tr[0] = t_lead0; tr[1] = t_stage0; tr[2] = t1; tr[3] = t1; }
      ctl->ns_leader += t1 - t_lead0; ctl->ns_stage += t1 - t_stage0;
      st_release_u32(&ctl->release, (unsigned)gi + 1u);
    }
    if (ga.hyper_moves) stage_store(S, ga, chain);                   // for the host and the kernels after the launch
    __syncthreads();
  }
}

// mhSpikeAllocation (Pr:528-638) for one all-zero, inactive gene of the persistent kernel.  Works on the gene's global-memory
// state as the main moves stored it.  XgLikelihood is not cached here: for an all-zero gene it is beta * A(y) out of the spike
// (no detected library: the normal part is empty) and 0 inside (P:117), so the current value comes from `lps`.
struct SpikeOutP { double y, lv, t; uint32_t f; bool finite_state, moved; };
template <int RT>
__device__ __noinline__ void spike_move_p(const FastH &H, const MathTabs &T, const FastArgs &fa, int chain, int64_t g, int64_t i, uint64_t step, SpikeOutP &o) {
  const SweepArgs &a = fa.s;
  uint32_t f = a.flags[i];
  double y = a.Yg[i], v = a.vg[i];
  o.moved = false; o.t = 0;
  if (!(f & F_ACTIVE)) {                                              // Pr:533: all-zero AND inactive
    const bool cur_in = (f & F_INSPIKE) != 0, prop_in = !cur_in;      // Pr:535
    uint32_t b2[4];
    double zy, zv;
    philox_block(a.seed, chain, step, 2, (uint64_t)(a.gene_offset + a.perm[g]), b2);
    box_muller(u32_open(b2[0]), u32_open(b2[1]), zy, zv);
    const double u = u32_open(b2[2]);
    const double yp = fma(H.sd_i, zy, H.im);                          // Pr:545
    const double lsp = fma(H.s1, yp, H.s0);                           // log(Sg_proposed), Pr:547
    const double lvp = (lsp + H.tau) + H.rtau * zv;                   // Pr:549: variance_g' = exp(.)
    const double vp = fexp(lvp, T.exp_t);
    const double PVp = vg_prior_from_q(H, lvp, vg_q(H, lvp, yp));     // Pr:556
    const double lv = flog(v, T.log_t);
    const double PVc = vg_prior_from_q(H, lv, vg_q(H, lv, y));        // Pr:558
    const double dp = yp - H.im, dc = y - H.im;
    const double pygp = prop_in ? H.log_sp : H.log_1msp - H.log_norm_i - dp * dp * H.inv2iv;   // Pr:560
    const double pyg = cur_in ? H.log_sp : H.log_1msp - H.log_norm_i - dc * dc * H.inv2iv;     // Pr:563
    const double pp = prop_in ? H.log_sp : (pygp + PVp);              // Pr:569-570
    const double pc = cur_in ? H.log_sp : (pyg + PVc);                // Pr:572-573
    double Ap = 0, tp = 0;
    if (!prop_in) { tp = a.gl[g] * fexp(yp, T.exp_t); Ap = detect_logsum<RT>(H, T, ~0ull, tp); }
    double Lp = prop_in ? 0.0 : times_beta(H, Ap);                    // Pr:579 (P:117 gives log(1*1) = 0 in the spike)
    double Lc = cur_in ? 0.0 : times_beta(H, fa.lps[i]);              // Pr:581: the cache, re-derived
    if (Lp == -INFINITY && Lc == -INFINITY) { Lp = 0; Lc = 0; }       // Pr:583-585
    const double HR = prop_in ? (pyg - H.log_1msp + PVc) : -(pygp - H.log_1msp + PVp);   // Pr:592-595
    const double ratio = H.temp * (Lp - Lc + pp - pc) + HR;           // Pr:601
    const int acc_s = (flog_core(u, T.log_t) < ratio);                // Pr:605
    if (isnan(ratio)) atomicOr(a.nan_flag, 1);
    if (acc_s) {
      f ^= F_INSPIKE;                                                 // Pr:616
      a.flags[i] = (uint16_t)f;
      if (!prop_in) { y = yp; v = vp; a.Yg[i] = yp; a.vg[i] = vp; fa.lps[i] = Ap; o.moved = true; o.t = tp; }   // Pr:620-636
    }
  }
  o.y = y; o.lv = (f & F_INSPIKE) ? 0.0 : flog(v, T.log_t); o.f = f;
  o.finite_state = isfinite(y) && isfinite(v);
}

// BLK threads per CTA (co-resident by cooperative launch).  Measured at 1 M x 16 (profiles/r2_summary.md): 3 CTAs x 256 threads at
// 80 registers 66.0 us per sweep; 1 x 768 (the same 24 warps, a third of the CTAs) 62.4; 1 x 512 at 128 registers 60.2; 1 x 640 at
// 96 registers 58.7; 576, 608, 672, 704 in between or worse.  Fewer, fatter warps win: a 32-gene tile takes a warp 7 us at 80
// registers and ~5 us at 96-128, so rounds are finer and the tail shorter.  launch_generations picks 640 for one chain over a
// large shard, 512 otherwise, and 3 x 256 for whole generations of batched chains (another chain's tiles fill the SM while the
// last CTA of a chain runs its scalar stage).  Results do not depend on the shape.
// The 256-thread shape only ever runs tuning generations (launch_generations): there the two accept windows, which are read where the
// tuning decision falls and come from DRAM (a burn-in generation touches 122 B per gene), are prefetched into L1 with the next tile's
// staged words — 264 -> 258.5 us per burn-in generation of 64 chains x 60 k x 4.  In the shapes that also run without tuning the same
// change costs those launches 0.3-0.6 us (register allocation) for 2 us per burn-in generation, and is left out.  The lines are
// only touched by this warp within a generation, and the generation's acquire orders the previous owner's stores before these loads.
constexpr int GEN_BLK_TUNED = 256;
template <int BLK> __device__ __forceinline__ void prefetch_windows(const SweepArgs &a, int64_t i) {
  if (BLK == GEN_BLK_TUNED && a.tune) { prefetch_l1(a.win_y + i); prefetch_l1(a.win_v + i); }
}

template <int RT, int KT, int BLK>
__global__ void __launch_bounds__(BLK, BLK <= 256 ? 3 : 1) k_generations(const __grid_constant__ GenArgs ga) {
  __shared__ FastH H;
  __shared__ MathTabs T;
  __shared__ unsigned sh_rel, sh_ticket, sh_qdone;
  extern __shared__ double dyn_smem[];                 // [nstat][BLOCK] int64 statistic columns, then [8 warps][10][32] staging
  const FastArgs &fa = ga.fa;
  const SweepArgs &a = fa.s;
  const int chain = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int K1 = a.K + 1, nstat = 3 * K1 + NSTAT_FIXED_;
  long long *acc = reinterpret_cast<long long *>(dyn_smem);
  GenCtl *ctl = ga.ctl + chain;
  long long *gacc = ga.acc + (int64_t)chain * nstat;
  for (int s = 0; s < nstat; ++s) acc[s * BLK + threadIdx.x] = 0;
  load_tabs(T, fa.exp_tab, fa.log_tab);
  // steward mode (one chain): the LAST CTA sweeps no tiles — it keeps the chain's scalar state in its shared memory for the whole
  // launch, prepares everything of the coming scalar stage that needs no statistics while the others sweep, and counts tickets
  const unsigned n_cta = gridDim.x - (ga.steward ? 1u : 0u);           // CTAs that sweep tiles
  if (ga.steward && blockIdx.x == n_cta) {
    __syncthreads();
    steward_loop<KT>(ga, T, reinterpret_cast<char *>(dyn_smem), ctl, gacc, nstat, chain, n_cta);
    return;
  }
  const int ntiles = (int)((a.g_end - a.g_begin + 31) / 32);
  // warp w of CTA b is global warp w * n_cta + b: consecutive tiles go to DIFFERENT CTAs, so a sweep with fewer tiles than warps
  // spreads over all SMs instead of filling the first CTAs (a tile runs faster the fewer warps share its SM)
  const int warp0 = wid * (int)n_cta + (int)blockIdx.x, nwarps = (int)n_cta * (BLK / 32);
  double *stg = dyn_smem + (int64_t)nstat * BLK + wid * 320 + lane;
  bool saw_nan = false;
  const uint64_t pol = l2_evict_first_policy();

  for (int gi = 0; gi < ga.ngen; ++gi) {
    const unsigned gen_no = (unsigned)gi;
    // ---- 1. this generation's hyper-parameters are published once release >= gen_no
    if (threadIdx.x == 0) {
      const unsigned *relp = &ctl->rel[(blockIdx.x % GEN_NREL) * 32];
      unsigned r = ld_acquire_u32(relp);
      if (r < gen_no) {
        const unsigned long long t0 = globaltimer_ns();
        unsigned ns = 32;
        while ((r = ld_acquire_u32(relp)) < gen_no) {
          if ((long long)(globaltimer_ns() - t0) > ga.timeout_ns) { r = GEN_ABORT; break; }
          __nanosleep(ns);
          if (ns < 256) ns += ns;                                       // back off: the scalar stage lasts microseconds
        }
      }
      sh_rel = r; sh_qdone = 0;
    }
    __syncthreads();
    if (sh_rel == GEN_ABORT) break;
    // per-CTA timeline (tools/gen_trace.py); compiled in only with -DZZ_GEN_TRACE_BUILD: two more live values in the tile loop cost spills
#ifdef ZZ_GEN_TRACE_BUILD
#define ZZ_TRACE(slot) do { if (ga.trace && threadIdx.x == 0 && gi < ga.trace_gens) ga.trace[((size_t)gi * gridDim.x + blockIdx.x) * 4 + (slot)] = globaltimer_ns(); } while (0)
#else
#define ZZ_TRACE(slot) ((void)0)
#endif
    ZZ_TRACE(0);
    {   // FastH was written by another SM: read it through L2
      const uint64_t *src = reinterpret_cast<const uint64_t *>(fa.fasth + chain);
      uint64_t *dst = reinterpret_cast<uint64_t *>(&H);
      for (int j = threadIdx.x; j < (int)(sizeof(FastH) / 8); j += blockDim.x) dst[j] = __ldcg(src + j);
      __syncthreads();
    }
    const uint64_t step = a.step + (uint64_t)ga.step_stride * (uint64_t)gi;
    const bool px_on = H.px_on != 0;
    ZZ_TRACE(1);

    // ---- 2. the sweep: tiles are taken from the END of the sorted gene order first (the all-zero genes with their extra move and
    // the weakly expressed genes with the full detection loop run while every warp is busy)
    uint32_t f_next = 0;
    // two static rounds when there are at least two FULL rounds of tiles; with fewer, one: a partial second round is dealt
    // dynamically, so the first warps to finish take it wherever they are — static, it would double the load of the first CTAs only
    // (125 k genes, the 8-GPU shard of the 1 M workload: 3 907 tiles on 3 544 warps; the 46 CTAs that owned the 363 extra tiles
    // finished at 19 us, everybody else at 8 us)
    const bool two_static = !ga.dyn_tiles || ntiles >= 2 * nwarps;
    const int n_static = two_static ? 2 * nwarps : nwarps, n_dyn = ntiles > n_static ? ntiles - n_static : 0;
    int tq = warp0 % GEN_NQ, tq_tries = 0;
    int t_ = warp0 < ntiles ? warp0 : -1;
    int t_n = (two_static && t_ >= 0 && warp0 + nwarps < ntiles) ? warp0 + nwarps : -1;
    unsigned ticket = 0;
    // one static round: every warp asks its home queue ONCE for a second tile, before it stages its first one (the answer arrives
    // while those copies are in flight).  n_dyn < nwarps, so queue q (tiles j = q mod GEN_NQ) has at least as many askers (warps
    // w = q mod GEN_NQ) as tiles: every tile is claimed, nobody needs to try a second queue
    const bool ask_first = !two_static && t_ >= 0 && n_dyn > 0;
    if (ask_first && lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&ctl->tile_q[tq * 32]) : "memory");
    if (t_ >= 0) {
      const int64_t g0 = a.g_begin + (int64_t)(ntiles - 1 - t_) * 32 + lane;
      if (g0 < a.g_end) { stage_tile_ef(a, fa, chain, g0, stg, pol); prefetch_windows<BLK>(a, (int64_t)chain * a.G + g0); f_next = a.flags[(int64_t)chain * a.G + g0]; }
    }
    if (ask_first) {
      const unsigned tk = __shfl_sync(0xffffffffu, ticket, 0);
      const int cnt = n_dyn > tq ? (n_dyn - tq + GEN_NQ - 1) / GEN_NQ : 0;
      if (tk < (unsigned)cnt) t_n = n_static + tq + GEN_NQ * (int)tk;
      tq_tries = GEN_NQ;                                                 // no further requests in this sweep
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    while (t_ >= 0) {
      const int64_t g = a.g_begin + (int64_t)(ntiles - 1 - t_) * 32 + lane;
      const int64_t i = (int64_t)chain * a.G + g;
      ZZ_DEV_ASSERT(t_ < ntiles && (t_n < 0 || (t_n < ntiles && t_n != t_)) && g >= a.g_begin);
      asm volatile("cp.async.wait_group 0;" ::: "memory");            // this lane's copies of the current tile have landed
      const bool valid = g < a.g_end;                                  // only the last (first processed) tile can be ragged
      const double y0 = stg[0 * 32], v0 = stg[1 * 32], ty = stg[3 * 32], tv = stg[4 * 32], gl = stg[5 * 32];
      double A0 = stg[2 * 32];
      const double xbar = stg[6 * 32], ss = stg[7 * 32];
      const uint64_t mask = (uint64_t)__double_as_longlong(stg[8 * 32]);
      const uint64_t gid = (uint64_t)(a.gene_offset + (int64_t)(uint32_t)__double2loint(stg[9 * 32]));
      uint32_t f = f_next;
      const int64_t g_n = a.g_begin + (int64_t)(ntiles - 1 - t_n) * 32 + lane;
      const bool next_valid = t_n >= 0 && g_n < a.g_end;
      if (next_valid) {                                                 // consumed one tile later (pinned here: volatile asm)
        unsigned short fw;
        asm volatile("ld.global.u16 %0, [%1];" : "=h"(fw) : "l"(a.flags + (int64_t)chain * a.G + g_n) : "memory");
        f_next = fw;
      }
      // every staged word of this tile is in registers: refill the slots with the warp's next tile, whose HBM latency then hides
      // behind this tile's ~1300 instructions (also for the lanes beyond the end of a ragged tile: their next tile is a full one)
      if (next_valid) { stage_tile_ef(a, fa, chain, g_n, stg, pol); prefetch_windows<BLK>(a, (int64_t)chain * a.G + g_n); }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // the tile after the next one: ask the home queue now, read the answer when this tile is done
      // (inline asm: the compiler would otherwise sink the atomic down to the shuffle that reads it and expose its whole latency)
      const bool asked = ga.dyn_tiles && t_n >= 0 && n_dyn > 0 && tq_tries < GEN_NQ;
      if (asked && lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&ctl->tile_q[tq * 32]) : "memory");
      if (valid) {
      const double nd_d = (double)((RT > 0 ? RT : H.R) - __popcll(mask));
      const bool live = !(f & F_INSPIKE);
      if (H.px_prev_acc && live) A0 += ga.dlt[i];                       // the accepted alpha_r of the previous generation (Pr:213-215, as an increment)
      uint32_t b0[4], b1[4];
      philox_block(a.seed, chain, step, 0, gid, b0);
      philox_block(a.seed, chain, step, 1, gid, b1);
      const double u_y = u32_to_open01(b0[2]), u1 = u32_to_open01(b0[3]), u2 = u32_to_open01(b1[0]);
      const double u_ai = u32_to_open01(b1[1]), u_w = u32_to_open01(b1[2]);
      const double z = fsqrt_nb(-2.0 * flog_nb(u32_to_open01(b0[0]), T.log_t)) * cospi(2.0 * u32_to_open01(b0[1]));
      const double lv = flog_nb(v0, T.log_t), rv = frcp_nb(v0);
      const bool bad = !is_pos_normal(v0);                              // poisoned state: reported through the NaN flag
      // ---- mhYg (Pr:250-333)
      const double yp = y0 + ty * z;                                     // Pr:257
      const double tp = gl * fexp_nb(yp, T.exp_t);                       // U:157
      const bool saturated = __all_sync(__activemask(), !live || tp * H.alpha_min >= 37.5);
      const double Ap = saturated ? (mask ? -INFINITY : 0.0)
                        : (RT > 0 && H.uclamp_ok) ? detect_logsum_sl<(RT > 0 && RT <= 18) ? RT : 1>(H, T, mask, tp) : detect_logsum_slow<RT>(H, T, mask, tp);
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
      const double v1 = acc_v ? vp : v0, lv1 = acc_v ? lv + arg : lv;
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
      // ---- stores (the two likelihood caches of the reference are not maintained inside a device-resident segment)
      a.flags[i] = (uint16_t)f;
      if (live) { __stcs(a.Yg + i, y1); __stcs(a.vg + i, v1); __stcs(fa.lps + i, A1); }   // streaming stores (L2 evict-first)
      if (a.tune && live) {                                              // Pr:312-322, Pr:40-50
        ulonglong2 w = BLK == GEN_BLK_TUNED ? a.win_y[i] : __ldcs(a.win_y + i); window_push(w, acc_y); __stcs(a.win_y + i, w);
        w = BLK == GEN_BLK_TUNED ? a.win_v[i] : __ldcs(a.win_v + i); window_push(w, acc_v); __stcs(a.win_v + i, w);
      }
      saw_nan |= live && (isnan(ratio_y) || isnan(ratio_v) || bad);
      double ys = y1, lvs = lv1;
      bool finite_state = isfinite(y1) && isfinite(v1);
      bool live_f = live;
      double t_fin = tp; bool t_known = acc_y;                           // gl * exp(final Yg), known when the proposal was accepted
      if ((f & F_ALLZERO) && fa.do_spike_move) {
        SpikeOutP so;
        spike_move_p<RT>(H, T, fa, chain, g, i, step, so);
        ys = so.y; lvs = so.lv; f = so.f; finite_state = so.finite_state; live_f = !(f & F_INSPIKE);
        if (so.moved) { t_fin = so.t; t_known = true; }
      }
      // ---- the gene's term of the pending mhP_x ratio: log q(alpha') - log q(alpha) of library px_lib at the final Yg (Pr:209-224)
      double pxd = 0; bool pxbad = false;
      if (px_on) {
        if (__any_sync(__activemask(), live_f && !t_known)) { const double tc = gl * fexp_nb(ys, T.exp_t); if (!t_known) t_fin = tc; }
        const bool px_sat = __all_sync(__activemask(), !live_f || t_fin * H.px_amin >= 37.5);
        const bool nd_lib = (((uint32_t)mask >> H.px_lib) & 1u) != 0;
        double d = 0;
        if (!px_sat) {
          const double tc = fmin(t_fin, H.t_clamp_px);
          const double qn = det_q(tc, H.px_am256, nd_lib, T), qc = det_q(tc, H.am256[H.px_lib], nd_lib, T);
          d = det_log_product(qn, T.log_t) - det_log_product(qc, T.log_t);
        } else if (nd_lib) d = NAN;                                      // log(0) - log(0): a state the chain cannot be in
        if (live_f) {
          if (H.beta == 0.0) d = 0;                                      // P:121-125
          pxbad = !isfinite(d);
          if (pxbad) d = 0;
          pxd = H.beta * d;
          ga.dlt[i] = d;                                                 // the increment of A(y) if the move is accepted
        }
      }
      // ---- sufficient statistics, fixed point
      {
        long long *col = acc + threadIdx.x;
        const bool act = f & F_ACTIVE, sp = f & F_INSPIKE;
        const int comp = act ? flag_k(f) : 0;                               // Pr:415 all_allocations
        ZZ_DEV_ASSERT(comp >= 0 && comp <= a.K && kn >= 1 && kn <= KT && gid - (uint64_t)a.gene_offset < (uint64_t)a.G);
        const long long fy = to_fx(ys), fy2 = to_fx(ys * ys);
        col[comp * BLK] += FX_ONE;
        if (act || !sp) { col[(K1 + comp) * BLK] += fy; col[(2 * K1 + comp) * BLK] += fy2; }
        long long *q = col + 3 * K1 * BLK;
        if (sp) q[0] += FX_ONE;
        else {
          q[1 * BLK] += FX_ONE; q[2 * BLK] += to_fx(lvs); q[3 * BLK] += to_fx(lvs * lvs); q[4 * BLK] += fy;
          q[5 * BLK] += fy2; q[6 * BLK] += to_fx(ys * lvs);
          if (px_on) q[FXS_PXDELTA * BLK] += to_fx(pxd);
        }
        if (pxbad) q[FXS_PXBAD * BLK] += FX_ONE;
        if (!finite_state) q[9 * BLK] += FX_ONE;
      }
      }   // valid
      int t_nn = -1;
      if (asked) {
        // The ticket asked for when this tile began.  Home queue exhausted: go round the others NOW, so that the next tile is known
        // before the one in hand starts and its loads hide behind it (deferring the search until the warp is out of tiles exposed
        // every later tile's load latency: +2.5 us per sweep, measured).  Queues found empty are remembered per CTA (sh_qdone), so at
        // the end of a sweep a CTA pays at most GEN_NQ failed round trips in all, not GEN_NQ per warp.
        unsigned tk = __shfl_sync(0xffffffffu, ticket, 0);
        for (;;) {
          const int cnt = n_dyn > tq ? (n_dyn - tq + GEN_NQ - 1) / GEN_NQ : 0;           // tiles in queue tq
          if (tk < (unsigned)cnt) { t_nn = n_static + tq + GEN_NQ * (int)tk; break; }
          if (lane == 0) atomicOr(&sh_qdone, 1u << tq);
          do { tq = (tq + 1) % GEN_NQ; ++tq_tries; } while (tq_tries < GEN_NQ && ((*(volatile unsigned *)&sh_qdone >> tq) & 1u));
          if (tq_tries >= GEN_NQ) break;
          if (lane == 0) ticket = atomicAdd(&ctl->tile_q[tq * 32], 1u);
          tk = __shfl_sync(0xffffffffu, ticket, 0);
        }
      } else if (!ga.dyn_tiles && t_n >= 0 && t_n + nwarps < ntiles) t_nn = t_n + nwarps;
      t_ = t_n; t_n = t_nn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // ---- 3. the CTA's column sums -> the chain's accumulators (integer atomics: order-free)
    __syncthreads();
    ZZ_TRACE(2);
    for (int s = wid; s < nstat; s += BLK / 32) {
      long long v = 0;
#pragma unroll
      for (int j = 0; j < BLK / 32; ++j) { v += acc[s * BLK + j * 32 + lane]; acc[s * BLK + j * 32 + lane] = 0; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v != 0) atomicAdd(reinterpret_cast<unsigned long long *>(gacc + s), (unsigned long long)v);
    }
    __syncthreads();
    // ---- 4. ticket.  Steward mode: done, the steward CTA counts the tickets.  Otherwise the last CTA of the chain to arrive runs the
    // scalar half of the generation.
    if (threadIdx.x == 0) { __threadfence(); sh_ticket = atomicAdd(&ctl->arrive, 1u); }   // cumulative: orders the CTA's reductions (barrier above) before the ticket
    ZZ_TRACE(3);
    if (ga.steward) continue;
    __syncthreads();
    if (sh_ticket != (gen_no + 1u) * n_cta - 1u) continue;
    __threadfence();
    const unsigned long long t_lead0 = globaltimer_ns();
    // the statistic columns are zero again: that memory is the stage's scratch (zeroed again below)
    const StageScr S = stage_scr(reinterpret_cast<char *>(dyn_smem));
    const bool ok = stage_totals(S, ga, ctl, gacc, nstat, chain, gi);
    if (!ok) break;
    const unsigned long long t_stage0 = globaltimer_ns();
    if (ga.hyper_moves) {
      stage_load(S, ga, chain);
      __syncthreads();
      stage_pre<KT>(S, T, a, step, step + (uint64_t)ga.step_stride);
      __syncthreads();
      stage_post<KT>(S, T, ga, ctl, step, gi == ga.ngen - 1, chain);
      stage_store(S, ga, chain);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x < GEN_NREL) st_release_u32(&ctl->rel[threadIdx.x * 32], gen_no + 1u);
    if (threadIdx.x == 0) {
      const unsigned long long t1 = globaltimer_ns();
      ctl->ns_leader += t1 - t_lead0; ctl->ns_stage += t1 - t_stage0;
      st_release_u32(&ctl->release, gen_no + 1u);
    }
    // the columns this CTA borrowed must be zero again before its next tile
    for (int j = threadIdx.x; j < (SCR_BYTES + 7) / 8; j += blockDim.x) dyn_smem[j] = 0.0;
    __syncthreads();
  }
  if (saw_nan) atomicOr(a.nan_flag, 1);
}

}  // namespace
