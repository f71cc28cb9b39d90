// =====================================================================================
// lgar_backward.cuh -- hand-written reverse-mode kernel (checkpointed recompute).
//
// dL/d(alpha, n, ksat)[L][B] for L = sum_{k,t,b} g[k][t][b] * out[k][t][b] (+ the same with the
// per-column sums), with the reference's autograd semantics.
//
// One WARP owns a tile of 32 columns and walks its forcing record backwards, chunk by chunk:
//   1. load the chunk-start checkpoint the forward kernel stored (lgar_forward(keep_checkpoints));
//   2. recompute the chunk with the plain fp64 sub-step (same code as the forward kernel), saving
//      the column state in front of EVERY sub-step into a per-warp ring in global memory;
//   3. for the sub-steps in reverse order: reload that state as tape leaves, re-run the sub-step
//      with R = Var (records the tape; takes exactly the forward branches because the values come
//      from the same cores), seed the adjoints of its outputs (dL/d out of this step + the adjoint
//      of the next state), sweep the tape backwards, keep the adjoint of the state in front of the
//      sub-step and accumulate the parameter adjoints.
// Tape, adjoints and ring live in global memory (L2), laid out lane-fastest so the 32 lanes of a
// warp touch consecutive addresses when they are at the same entry.  The root finders are
// straight-through (they never appear on the tape); Geff is one macro entry whose partials are
// evaluated cooperatively by the warp.
// =====================================================================================
#pragma once
#include "lgar_forward.cuh"

namespace lgar {

constexpr int NPAR_IDS = 3 * MAXL;  // alpha[l] = 3l, n[l] = 3l+1, ksat[l] = 3l+2
template <int FM>
__host__ __device__ constexpr int leaf_fields() { return NPAR_IDS; }                  // + fld*FM + i
template <int FM>
__host__ __device__ constexpr int leaf_ponded() { return NPAR_IDS + 5 * FM; }
template <int FM>
__host__ __device__ constexpr int leaf_endvol() { return NPAR_IDS + 5 * FM + 1; }
template <int FM>
__host__ __device__ constexpr int leaf_giuh() { return NPAR_IDS + 5 * FM + 2; }
template <int FM>
__host__ __device__ constexpr int num_leaves() { return NPAR_IDS + 5 * FM + 2 + NGIUH; }
template <int FM>
__host__ __device__ constexpr int ring_doubles() { return 5 * FM + S_SUMS; }  // state record without the sums

struct BParams {
  KParams K;
  const double* grad_per_step;  // [popc(grad_mask)][T][B] or NULL
  uint32_t grad_mask;
  const double* grad_sums;      // [NOUT][B] or NULL
  double* grad_alpha;           // [L][B]
  double* grad_n;
  double* grad_ksat;
  // per resident warp scratch
  double* ring_d;               // [slots][ring_steps][ring_doubles][32]
  int32_t* ring_i;              // [slots][ring_steps][2][32]  (n, cntpk)
  uint8_t* ring_f;              // [slots][ring_steps][FM][32]
  TapeEntry* tape;              // [slots][tape_cap][32]
  double* adj;                  // [slots][num_leaves + tape_cap][32]
  double* lam;                  // [slots][num_leaves][32]
  unsigned long long* next_tile;
  int32_t ring_steps;           // chunk_steps * S
  int32_t tape_cap;
  int32_t* tape_overflow;       // [B] set to 1 if a column's tape overflowed (gradient invalid)
};

__host__ inline size_t backward_scratch_bytes(int /*B*/, int /*Bp*/, int /*L*/, int S, int FM, int chunk, int slots,
                                              int tape_cap) {
  const size_t ring_steps = (size_t)chunk * S;
  const size_t nd = 5 * (size_t)FM + S_SUMS, nl = NPAR_IDS + 5 * (size_t)FM + 2 + NGIUH;
  size_t per = ring_steps * nd * 32 * 8 + ring_steps * 2 * 32 * 4 + ring_steps * (size_t)FM * 32 +
               (size_t)tape_cap * 32 * sizeof(TapeEntry) + (nl + tape_cap) * 32 * 8 + nl * 32 * 8;
  return per * slots + 4096;
}

template <int FM, class R>
__device__ void ring_save(const BParams& P, int slot, int j, int lane, Tile<FM, R>& T) {
  Column<FM, R>& C = T.col;
  double* rd = P.ring_d + (((size_t)slot * P.ring_steps + j) * ring_doubles<FM>()) * 32 + lane;
  for (int i = 0; i < C.n; i++)
#pragma unroll
    for (int k = 0; k < 5; k++) rd[(size_t)(k * FM + i) * 32] = C.f(k, i);
  double* rs = rd + (size_t)(5 * FM) * 32;
  rs[S_PONDED * 32] = val(C.ponded_water);
  rs[S_PREV_PRECIP * 32] = C.previous_precip;
  rs[S_END_VOL * 32] = val(C.ending_volume);
  for (int i = 0; i < NGIUH; i++) rs[(S_GIUH + i) * 32] = val(C.giuh[i]);
  int32_t* ri = P.ring_i + (((size_t)slot * P.ring_steps + j) * 2) * 32 + lane;
  ri[0] = C.n;
  ri[32] = (int32_t)C.cntpk;
  uint8_t* rf = P.ring_f + (((size_t)slot * P.ring_steps + j) * FM) * 32 + lane;
  for (int i = 0; i < C.n; i++) rf[(size_t)i * 32] = C.gb[i * NT];
}
template <int FM>
__device__ void ring_load_as_leaves(const BParams& P, int slot, int j, int lane, Tile<FM, Var>& T) {
  Column<FM, Var>& C = T.col;
  const double* rd = P.ring_d + (((size_t)slot * P.ring_steps + j) * ring_doubles<FM>()) * 32 + lane;
  const int32_t* ri = P.ring_i + (((size_t)slot * P.ring_steps + j) * 2) * 32 + lane;
  C.n = ri[0];
  C.cntpk = (unsigned)ri[32];
  for (int i = 0; i < C.n; i++)
#pragma unroll
    for (int k = 0; k < 5; k++) C.s(k, i, Var(rd[(size_t)(k * FM + i) * 32], leaf_fields<FM>() + k * FM + i));
  const double* rs = rd + (size_t)(5 * FM) * 32;
  C.ponded_water = Var(rs[S_PONDED * 32], leaf_ponded<FM>());
  C.previous_precip = rs[S_PREV_PRECIP * 32];
  C.ending_volume = Var(rs[S_END_VOL * 32], leaf_endvol<FM>());
  for (int i = 0; i < NGIUH; i++) C.giuh[i] = Var(rs[(S_GIUH + i) * 32], leaf_giuh<FM>() + i);
  const uint8_t* rf = P.ring_f + (((size_t)slot * P.ring_steps + j) * FM) * 32 + lane;
  for (int i = 0; i < C.n; i++) C.gb[i * NT] = rf[(size_t)i * 32];
}

template <int FM>
__global__ void __launch_bounds__(NT) lgar_backward_kernel(const BParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm_fields = reinterpret_cast<double*>(smem_raw);                       // [5*FM][NT]
  double* sm_nodes = sm_fields + 5 * FM * NT;                                    // [WARPS][NODEBUF]
  short* sm_ids = reinterpret_cast<short*>(sm_nodes + WARPS * NODEBUF);          // [5*FM][NT]
  uint8_t* sm_flags = reinterpret_cast<uint8_t*>(sm_ids + 5 * FM * NT);          // [FM][NT]
  __shared__ unsigned long long sm_item[WARPS];

  const KParams& K = P.K;
  const lgar_problem& p = K.p;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * WARPS + warp;  // resident-warp scratch slot
  double* nodebuf = sm_nodes + warp * NODEBUF;
  const int Tn = p.num_steps, S = p.num_subcycles;
  const size_t B = p.num_columns;
  constexpr int NL = num_leaves<FM>();

  TapeEntry* tape = P.tape + (size_t)slot * P.tape_cap * 32 + lane;
  double* adj = P.adj + (size_t)slot * (NL + P.tape_cap) * 32 + lane;
  double* lam = P.lam + (size_t)slot * NL * 32 + lane;
  TapeCtl& tc = g_tapectl[threadIdx.x];
  tc.base = tape;
  tc.cap = P.tape_cap;
  tc.first_id = NL;

  Tile<FM, double> Td;
  Td.col.fb = sm_fields + threadIdx.x;
  Td.col.ib = nullptr;
  Td.col.gb = sm_flags + threadIdx.x;
  Td.ctx.iter_cap = K.iter_cap;
  Tile<FM, Var> Tv;
  Tv.col.fb = sm_fields + threadIdx.x;
  Tv.col.ib = sm_ids + threadIdx.x;
  Tv.col.gb = sm_flags + threadIdx.x;
  Tv.ctx.iter_cap = K.iter_cap;

  for (;;) {
    if (lane == 0) sm_item[warp] = atomicAdd(P.next_tile, 1ULL);
    __syncwarp();
    const unsigned long long tile = sm_item[warp];
    __syncwarp();
    if (tile >= (unsigned long long)K.ntiles) break;
    const int b = (int)tile * 32 + lane;
    const bool valid = (size_t)b < B;
    const int bb = valid ? b : (int)B - 1;

    // final status of the forward pass: steps at and after the crash step have no gradient
    const int final_st = __ldcg(K.state_i + ((size_t)K.nchunks * NI_STATE + 2) * K.Bp + bb);
    const int final_crash = __ldcg(K.state_i + ((size_t)K.nchunks * NI_STATE + 3) * K.Bp + bb);
    const int t_end = (final_st == 0) ? Tn : final_crash;  // steps [0, t_end) produced outputs

    load_params(K, bb, Td);
    load_params(K, bb, Tv);
    for (int l = 0; l < Tv.col.L; l++) {
      Tv.col.soil[l].id_alpha = 3 * l;
      Tv.col.soil[l].id_n = 3 * l + 1;
      Tv.col.soil[l].id_ksat = 3 * l + 2;
      Tv.col.soil[l].id_m = -1;
    }
    const int site = (p.site_index && valid) ? __ldg(p.site_index + b) : 0;
    const double* frc = p.forcing + (size_t)site * Tn * 2;
    for (int q = 0; q < NL; q++) lam[(size_t)q * 32] = 0.0;
    double gpar[NPAR_IDS];
#pragma unroll
    for (int q = 0; q < NPAR_IDS; q++) gpar[q] = 0.0;
    bool overflow = false;
    // gradient of the initial state w.r.t. the parameters is added after the loop (theta_init(alpha, n))

    for (int chunk = K.nchunks - 1; chunk >= 0; chunk--) {
      const int t0 = chunk * K.chunk_steps;
      const int t1 = min(Tn, t0 + K.chunk_steps);
      // ---- 1+2: recompute the chunk, saving the state in front of every sub-step
      load_state(K, chunk, bb, Td);
      Td.ctx.st = __ldcg(K.state_i + ((size_t)chunk * NI_STATE + 2) * K.Bp + bb);
      precompute_psi_wp(Td, p.wilting_point_psi);
      for (int t = t0; t < t1; t++) {
        const double2 x = __ldg(reinterpret_cast<const double2*>(frc) + t);
        const bool alive = valid && (Td.ctx.st == 0) && (t < t_end);
        for (int sc = 0; sc < S; sc++) {
          ring_save(P, slot, (t - t0) * S + sc, lane, Td);
#pragma unroll
          for (int k = 0; k < NOUT; k++) Td.acc[k] = 0.0;
          substep(Td, alive, x.x, x.y, K, nodebuf);
        }
      }
      __syncwarp();
      // ---- 3: taped replay in reverse
      for (int t = t1 - 1; t >= t0; t--) {
        const double2 x = __ldg(reinterpret_cast<const double2*>(frc) + t);
        const bool alive = valid && (t < t_end);
        for (int sc = S - 1; sc >= 0; sc--) {
          const int j = (t - t0) * S + sc;
          tc.n = 0;
          Tv.ctx.st = 0;
          ring_load_as_leaves(P, slot, j, lane, Tv);
          // derived parameters on the tape: m = 1 - 1/n (data/utils.py:75), psi_wp (aet.py:37-43)
          for (int l = 0; l < Tv.col.L; l++) {
            SoilT<Var>& s = Tv.col.soil[l];
            Var mV = 1.0 - (1.0 / Var(s.n, s.id_n));
            s.id_m = mV.id;
          }
          precompute_psi_wp(Tv, p.wilting_point_psi);
#pragma unroll
          for (int k = 0; k < NOUT; k++) Tv.acc[k] = Var(0.0);
          substep(Tv, alive, x.x, x.y, K, nodebuf);
          __syncwarp();
          const int ne = min(tc.n, tc.cap);
          if (tc.n > tc.cap) overflow = true;
          if (alive && Tv.ctx.st == 0) {
            // zero the adjoints, then seed
            for (int q = 0; q < NL + ne; q++) adj[(size_t)q * 32] = 0.0;
            Column<FM, Var>& C = Tv.col;
            for (int i = 0; i < C.n; i++)
#pragma unroll
              for (int k = 0; k < 5; k++) {
                const int id = C.fid(k, i);
                if (id >= 0) adj[(size_t)id * 32] += lam[(size_t)(leaf_fields<FM>() + k * FM + i) * 32];
              }
            if (C.ponded_water.id >= 0) adj[(size_t)C.ponded_water.id * 32] += lam[(size_t)leaf_ponded<FM>() * 32];
            if (C.ending_volume.id >= 0) adj[(size_t)C.ending_volume.id * 32] += lam[(size_t)leaf_endvol<FM>() * 32];
            for (int i = 0; i < NGIUH; i++)
              if (C.giuh[i].id >= 0) adj[(size_t)C.giuh[i].id * 32] += lam[(size_t)(leaf_giuh<FM>() + i) * 32];
            // dL / d(outputs of this forcing step)
#pragma unroll
            for (int k = 0; k < NOUT; k++) {
              double g = 0.0;
              const bool state_out = (k == LGAR_OUT_ENDING_VOLUME || k == LGAR_OUT_PONDED_WATER);
              if (P.grad_per_step && (P.grad_mask & (1u << k)))
                g += __ldg(P.grad_per_step + ((size_t)__popc(P.grad_mask & ((1u << k) - 1u)) * Tn + t) * B + b);
              if (P.grad_sums && (!state_out || t == Tn - 1)) g += __ldg(P.grad_sums + (size_t)k * B + b);
              if (g == 0.0) continue;
              int id = -1;
              if (k == LGAR_OUT_ENDING_VOLUME) id = (sc == S - 1) ? C.ending_volume.id : -1;
              else if (k == LGAR_OUT_PONDED_WATER) id = (sc == S - 1) ? C.ponded_water.id : -1;
              else id = Tv.acc[k].id;
              if (id >= 0) adj[(size_t)id * 32] += g;
            }
            // reverse sweep
            for (int e = ne - 1; e >= 0; e--) {
              const double g = adj[(size_t)(NL + e) * 32];
              if (g != 0.0) {
                const TapeEntry te = tape[(size_t)e * 32];
                if (te.a >= 0) adj[(size_t)te.a * 32] += g * te.da;
                if (te.b >= 0) adj[(size_t)te.b * 32] += g * te.db;
              }
            }
            for (int q = NPAR_IDS; q < NL; q++) lam[(size_t)q * 32] = adj[(size_t)q * 32];
#pragma unroll
            for (int q = 0; q < NPAR_IDS; q++) gpar[q] += adj[(size_t)q * 32];
          }
          __syncwarp();
        }
      }
    }
    // ---- the initial state depends on the parameters: theta_init = theta_l(psi_init), K_init
    //      (data/utils.py:82-84, WettingFront.py:38-48); ending_volume(0) = mass_balance()
    {
      tc.n = 0;
      Tv.ctx.st = 0;
      for (int l = 0; l < Tv.col.L; l++) {
        SoilT<Var>& s = Tv.col.soil[l];
        Var mV = 1.0 - (1.0 / Var(s.n, s.id_n));
        s.id_m = mV.id;
      }
      init_column(Tv, __ldg(p.initial_psi + bb));
      const int ne = min(tc.n, tc.cap);
      if (valid && Tv.ctx.st == 0) {
        for (int q = 0; q < NL + ne; q++) adj[(size_t)q * 32] = 0.0;
        Column<FM, Var>& C = Tv.col;
        for (int i = 0; i < C.n; i++)
#pragma unroll
          for (int k = 0; k < 5; k++) {
            const int id = C.fid(k, i);
            if (id >= 0) adj[(size_t)id * 32] += lam[(size_t)(leaf_fields<FM>() + k * FM + i) * 32];
          }
        if (C.ending_volume.id >= 0) adj[(size_t)C.ending_volume.id * 32] += lam[(size_t)leaf_endvol<FM>() * 32];
        for (int e = ne - 1; e >= 0; e--) {
          const double g = adj[(size_t)(NL + e) * 32];
          if (g != 0.0) {
            const TapeEntry te = tape[(size_t)e * 32];
            if (te.a >= 0) adj[(size_t)te.a * 32] += g * te.da;
            if (te.b >= 0) adj[(size_t)te.b * 32] += g * te.db;
          }
        }
#pragma unroll
        for (int q = 0; q < NPAR_IDS; q++) gpar[q] += adj[(size_t)q * 32];
      }
    }
    if (valid) {
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      for (int l = 0; l < p.num_layers; l++) {
        P.grad_alpha[(size_t)l * B + b] = overflow ? qnan : gpar[3 * l];
        P.grad_n[(size_t)l * B + b] = overflow ? qnan : gpar[3 * l + 1];
        P.grad_ksat[(size_t)l * B + b] = overflow ? qnan : gpar[3 * l + 2];
      }
      if (P.tape_overflow) P.tape_overflow[b] = overflow ? 1 : 0;
    }
    __syncwarp();
  }
}

}  // namespace lgar
