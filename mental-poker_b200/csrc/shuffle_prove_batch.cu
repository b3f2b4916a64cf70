// shuffle_and_remask for batches (BASELINE config: batch of independent 52-card proofs): the
// lockstep small-deck prover and the worker-context dispatcher for large decks.
#include <chrono>

#include "shuffle_internal.cuh"

namespace mp {

// ------------------------------------------------------------------------------------------
// Lockstep batched prover for small decks (BASELINE config: batch of 52-card proofs).
//
// A 52-card proof is a chain of five dependent MSM launch sequences whose cost is latency, not
// work, so B proofs advance TOGETHER: every Fiat-Shamir round is one batched MSM launch over the
// jobs of all proofs (fixed-base table mode for every commitment, variable-base for the diagonal
// ciphertext MSMs).  At these sizes the scalar-field work is O(m^2 n) ~ a few thousand
// multiplications per proof, so it runs on the host threads that also own the transcripts (the
// device kernels of frvec.cu are for the 2^16-card path).  Results are byte-identical to
// mp_shuffle_and_remask (tests/test_gpu_shuffle.py).
// ------------------------------------------------------------------------------------------
namespace {

struct ProverHost {  // one proof's host-side state across the rounds
  Transcript fs;
  std::vector<fr> r, s, a, b, t, d, Bv, col, bk, sv;
  std::vector<fr> z_a0, z_bm1, z_t, sv_d, sv_delta, me_a0, me_b, me_s, me_tau;
  std::vector<fr> Az, Bz;  // zero-argument rows, (m + 1) x n each
  fr s_prod, z_r0, z_sm1, sv_rd, sv_s1, sv_sx, me_r0;
  fr x, y, z, xh, yh;
  std::vector<fr> xhp;
};

inline void put_fr(uint32_t* dst, const fr& v) { fr_to_canonical(v, dst); }

// E[(p*2m + k)*2 + comp] += enc part of proof p:  comp 0 -> g1[p*stride + c1_off + k], comp 1 -> c2_off + k
__global__ void __launch_bounds__(64) k_combine_E_batch(xyzz* __restrict__ E, const xyzz* __restrict__ g1, uint32_t stride,
                                                        uint32_t c1_off, uint32_t c2_off, int two_m, uint64_t total) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t p = g / (2 * (uint64_t)two_m);
  uint32_t rem = (uint32_t)(g % (2 * (uint64_t)two_m));
  uint32_t k = rem >> 1, comp = rem & 1;
  xyzz x = E[g], y = g1[p * stride + (comp ? c2_off : c1_off) + k];
  xyzz_add(x, y);
  E[g] = x;
}

}  // namespace

int32_t prove_sub_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                               const uint8_t* rhos, const uint8_t* rands, size_t Bs, uint8_t* out_decks, uint8_t* proofs,
                               int threads) {
  NvtxRange nvtx("prove_sub_batch");
  ShuffleState* S = ctx->shuffle;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n, plen = shuffle_proof_len(m, n), rlen = shuffle_randomness_len(m, n) * 32;
  const Layout L(m, n);
  cudaStream_t st = ctx->stream;
  const uint32_t nb = (uint32_t)(n + 4);       // bases of the fixed-base table: h, g_1..g_n, enc_g, ghat, pk
  const uint32_t tot = (uint32_t)(n + 1);      // scalars of one row commitment: blind + n values
  const int R = m + 4;                         // row commitments of round C
  const uint32_t SC = (uint32_t)(R * tot + 10 * m);   // round-C scalars per proof
  const uint32_t JC = (uint32_t)(R + 6 * m);          // round-C G1 jobs per proof
  const uint32_t SD = (uint32_t)(2 * tot + 2 * (2 * m + 1));  // round-D scalars per proof
  const uint32_t JD = (uint32_t)(2 * m + 3);

  // ---- round 0: remask every deck with one launch (the batch is one deck of Bs*N cards)
  {
    std::vector<uint32_t> gperm(Bs * N);
    for (size_t p = 0; p < Bs; p++)
      for (size_t i = 0; i < N; i++) {
        if (perms[p * N + i] >= N) return ctx->fail(MP_ERR_INVALID_ARG, "proof %zu: permutation entry %zu out of range", p, i);
        gperm[p * N + i] = (uint32_t)(p * N) + perms[p * N + i];
      }
    const void* d_shuffled = nullptr;
    int32_t rc = shuffle_remask(ctx, pk, decks, gperm.data(), rhos, Bs * N, out_decks, nullptr, &d_shuffled);
    if (rc != MP_OK) return rc;
  }
  int launches = ctx->launches;
  // MP_TRACE_BATCH=1: wall time of the host phases (transcripts + scalar algebra on `threads` threads) and of the
  // device phases (upload, MSM launches, download, synchronise) of this sub-batch, to stderr
  static const bool trace_on = [] { const char* e = getenv("MP_TRACE_BATCH"); return e && atoi(e) != 0; }();
  double t_host = 0, t_dev = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  auto t_mark = now();
  auto host_done = [&] { t_host += since(t_mark); t_mark = now(); };
  auto dev_done = [&] { t_dev += since(t_mark); t_mark = now(); };

  // ---- device buffers
  const size_t rows_max = std::max<size_t>(SC, (size_t)m * tot);
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, (Bs * N + 2) * kCtBytes);   // holds the remasked decks already
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, (Bs * N + 2) * 2 * sizeof(affine));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, Bs * (N + n) * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, Bs * 4 * (size_t)m * sizeof(xyzz));
  uint32_t* d_g1_scal = (uint32_t*)ctx->scratch(sG1Scal, Bs * rows_max * 32 + 64);
  xyzz* d_g1_out = (xyzz*)ctx->scratch(sG1Out, Bs * JC * sizeof(xyzz));
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sCanonOut, Bs * (JC + 4 * (size_t)m) * kPointBytes);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_ct_scal); NEED(d_ct_out); NEED(d_g1_scal); NEED(d_g1_out); NEED(d_canon); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  // pk column of the fixed-base table
  if (!S->ck_pk_valid || memcmp(S->ck_pk, pk, kPointBytes) != 0) {
    uint8_t* d_pk = (uint8_t*)ctx->scratch(sSmallUp, 256);
    NEED(d_pk);
    CK(cudaMemcpyAsync(d_pk, pk, kPointBytes, cudaMemcpyHostToDevice, st));
    CK(points_to_mont((const uint32_t*)d_pk, S->d_ck + (n + 3), 1, d_bad, st));
    CK(msm_build_table(ctx->ws, S->d_ck, nb, (uint32_t)(n + 3), 1, S->tab_c, S->d_tab_ck, st));
    memcpy(S->ck_pk, pk, kPointBytes);
    S->ck_pk_valid = true;
    launches += 3;
  }
  // the remasked decks (still in d_ct_canon) in Montgomery form for the diagonal MSMs
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, Bs * N * 2, nullptr, st));
  launches += 1;

  std::vector<ProverHost> H(Bs);
  std::vector<uint32_t> h_scal;
  std::vector<uint8_t> h_pts;
  std::vector<MsmJob> jobs;

  auto run_commit_jobs = [&](size_t n_scalars, size_t njobs_total, size_t npoints_out) -> int32_t {
    host_done();
    CK(cudaMemcpyAsync(d_g1_scal, h_scal.data(), n_scalars * 32, cudaMemcpyHostToDevice, st));
    CK(msm_run(ctx->ws, d_g1_scal, n_scalars, S->d_tab_ck, 1, jobs.data(), (int)njobs_total, S->tab_c, d_g1_out, st, 0, -1, nb));
    launches += msm_last_launches(ctx->ws);
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, npoints_out, st));
    launches += 1;
    h_pts.resize(npoints_out * kPointBytes);
    CK(cudaMemcpyAsync(h_pts.data(), d_canon, npoints_out * kPointBytes, cudaMemcpyDeviceToHost, st));
    CK(stream_wait(ctx, st));
    dev_done();
    return MP_OK;
  };
  int32_t rc;

  // ---- round A: c_A[k] = com(chunk_k(a); r_k)
  h_scal.assign(Bs * (size_t)m * tot * 8, 0);
  S->pool.run(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    RandCursor rcur{rands + p * rlen};
    h.r = rcur.vec(m);
    h.s = rcur.vec(m);
    h.s_prod = rcur.one();
    h.sv.assign((size_t)m, fr_zero());
    for (int i = 1; i < m - 1; i++) h.sv[i] = rcur.one();
    h.z_a0 = rcur.vec(n); h.z_bm1 = rcur.vec(n);
    h.z_r0 = rcur.one(); h.z_sm1 = rcur.one();
    h.z_t.assign((size_t)2 * m + 1, fr_zero());
    for (int k = 0; k <= 2 * m; k++) if (k != m + 1) h.z_t[k] = rcur.one();
    h.sv_d = rcur.vec(n);
    h.sv_rd = rcur.one();
    h.sv_delta.assign((size_t)n, fr_zero());
    h.sv_delta[0] = h.sv_d[0];
    for (int i = 1; i < n - 1; i++) h.sv_delta[i] = rcur.one();
    h.sv_s1 = rcur.one(); h.sv_sx = rcur.one();
    h.me_a0 = rcur.vec(n);
    h.me_r0 = rcur.one();
    h.me_b.assign((size_t)2 * m, fr_zero()); h.me_s = h.me_b; h.me_tau = h.me_b;
    for (int k = 0; k < 2 * m; k++) if (k != m) { h.me_b[k] = rcur.one(); h.me_s[k] = rcur.one(); h.me_tau[k] = rcur.one(); }
    h.a.resize(N);
    for (size_t i = 0; i < N; i++) h.a[i] = fr_from_u64((uint64_t)perms[p * N + i] + 1);
    for (int k = 0; k < m; k++) {
      uint32_t* dst = &h_scal[((p * m + k) * (size_t)tot) * 8];
      put_fr(dst, h.r[k]);
      for (int j = 0; j < n; j++) put_fr(dst + 8 * (size_t)(1 + j), h.a[(size_t)k * n + j]);
    }
  });
  jobs.assign(Bs * (size_t)m, MsmJob{});
  for (size_t q = 0; q < Bs * (size_t)m; q++) jobs[q] = MsmJob{(uint32_t)(q * tot), 0, tot};
  if ((rc = run_commit_jobs(Bs * (size_t)m * tot, Bs * (size_t)m, Bs * (size_t)m)) != MP_OK) return rc;

  // ---- round B: x; b_i = x^{perm[i]+1}; c_B[k] = com(chunk_k(b); s_k)
  S->pool.run(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    memcpy(proof + L.cA, &h_pts[p * (size_t)m * kPointBytes], (size_t)m * kPointBytes);
    absorb_statement(h.fs, S, pk, decks + p * N * kCtBytes, out_decks + p * N * kCtBytes, N, proof + L.cA);
    h.x = h.fs.challenge();
    std::vector<fr> xp = h_powers(h.x, (int)N + 1);
    h.b.resize(N);
    for (size_t i = 0; i < N; i++) h.b[i] = xp[perms[p * N + i] + 1];
  });
  // (h_scal is shared: fill it after the transcripts so round A's data is no longer needed)
  S->pool.run(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    for (int k = 0; k < m; k++) {
      uint32_t* dst = &h_scal[((p * m + k) * (size_t)tot) * 8];
      put_fr(dst, h.s[k]);
      for (int j = 0; j < n; j++) put_fr(dst + 8 * (size_t)(1 + j), h.b[(size_t)k * n + j]);
    }
  });
  if ((rc = run_commit_jobs(Bs * (size_t)m * tot, Bs * (size_t)m, Bs * (size_t)m)) != MP_OK) return rc;

  // ---- round C: y, z; product-argument rows, SVP and multi-exp first messages, diagonal ciphertexts
  h_scal.assign(Bs * (size_t)SC * 8, 0);
  std::vector<uint32_t> h_ct_scal(Bs * (N + n) * 8);
  S->pool.run(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    memcpy(proof + L.cB, &h_pts[p * (size_t)m * kPointBytes], (size_t)m * kPointBytes);
    h.fs.begin(); h.fs.feed_label("shuffle_argument_b"); h.fs.feed_points64(proof + L.cB, m); h.fs.end();
    h.y = h.fs.challenge();
    h.z = h.fs.challenge();
    h.d.resize(N); h.t.resize(m);
    for (size_t i = 0; i < N; i++) h.d[i] = fr_sub(fr_add(fr_mul(h.y, h.a[i]), h.b[i]), h.z);
    for (int k = 0; k < m; k++) h.t[k] = fr_add(fr_mul(h.y, h.r[k]), h.s[k]);
    h.Bv.resize(N);
    for (int j = 0; j < n; j++) {
      fr acc = h.d[j];
      h.Bv[j] = acc;
      for (int k = 1; k < m; k++) { acc = fr_mul(acc, h.d[(size_t)k * n + j]); h.Bv[(size_t)k * n + j] = acc; }
    }
    h.col.assign(h.Bv.begin() + (size_t)(m - 1) * n, h.Bv.end());
    h.bk.resize(n);
    h.bk[0] = h.col[0];
    for (int i = 1; i < n; i++) h.bk[i] = fr_mul(h.bk[i - 1], h.col[i]);
    h.sv[0] = h.t[0];
    h.sv[m - 1] = h.s_prod;
    fr rho_star = fr_zero();
    for (size_t i = 0; i < N; i++) rho_star = fr_sub(rho_star, fr_mul(h_fr(rhos + (p * N + i) * 32), h.b[i]));
    h.me_tau[m] = rho_star;
    // G1 scalars of this proof: R rows of (n + 1), then 2m pairs (s_k, b_k), 2m singles tau_k, 2m pairs (b_k, tau_k)
    uint32_t* base = &h_scal[p * (size_t)SC * 8];
    auto row = [&](int rix, const fr& blind, const fr* vals, int len) {
      uint32_t* dst = base + (size_t)rix * tot * 8;
      put_fr(dst, blind);
      for (int j = 0; j < len; j++) put_fr(dst + 8 * (size_t)(1 + j), vals[j]);
    };
    for (int i = 0; i < m; i++) row(i, h.sv[i], &h.Bv[(size_t)i * n], n);
    row(m, h.sv_rd, h.sv_d.data(), n);
    std::vector<fr> v1((size_t)n - 1), v2((size_t)n - 1);
    for (int i = 0; i + 1 < n; i++) {
      v1[i] = fr_neg(fr_mul(h.sv_delta[i], h.sv_d[i + 1]));
      v2[i] = fr_sub(fr_sub(h.sv_delta[i + 1], fr_mul(h.col[i + 1], h.sv_delta[i])), fr_mul(h.bk[i], h.sv_d[i + 1]));
    }
    row(m + 1, h.sv_s1, v1.data(), n - 1);
    row(m + 2, h.sv_sx, v2.data(), n - 1);
    row(m + 3, h.me_r0, h.me_a0.data(), n);
    uint32_t* sm = base + (size_t)R * tot * 8;
    for (int k = 0; k < 2 * m; k++) {
      put_fr(sm + 8 * (size_t)(2 * k), h.me_s[k]);
      put_fr(sm + 8 * (size_t)(2 * k + 1), h.me_b[k]);
      put_fr(sm + 8 * (size_t)(4 * m + k), h.me_tau[k]);
      put_fr(sm + 8 * (size_t)(6 * m + 2 * k), h.me_b[k]);
      put_fr(sm + 8 * (size_t)(6 * m + 2 * k + 1), h.me_tau[k]);
    }
    // ciphertext scalars: rows a0 | b_1..b_m
    uint32_t* cs = &h_ct_scal[p * (N + n) * 8];
    for (int j = 0; j < n; j++) put_fr(cs + 8 * (size_t)j, h.me_a0[j]);
    for (size_t i = 0; i < N; i++) put_fr(cs + 8 * ((size_t)n + i), h.b[i]);
  });
  // diagonal ciphertext MSMs of every proof: one variable-base launch sequence
  {
    std::vector<MsmJob> diag(Bs * 2 * (size_t)m);
    for (size_t p = 0; p < Bs; p++)
      for (int k = 0; k < 2 * m; k++) {
        int i0 = std::max(1, m - k), i1 = std::min(m, 2 * m - k);
        diag[p * 2 * m + k] = MsmJob{(uint32_t)(p * (N + n) + (size_t)(k - m + i0) * n), (uint32_t)(p * N + (size_t)(i0 - 1) * n),
                                     (uint32_t)((size_t)(i1 - i0 + 1) * n)};
      }
    host_done();
    CK(cudaMemcpyAsync(d_ct_scal, h_ct_scal.data(), h_ct_scal.size() * 4, cudaMemcpyHostToDevice, st));
    CK(msm_run(ctx->ws, d_ct_scal, Bs * (N + n), d_ct_mont, 2, diag.data(), (int)diag.size(), msm_pick_window(N / 2 + 1, diag.size()), d_ct_out, st));
    launches += msm_last_launches(ctx->ws);
  }
  jobs.clear();
  for (size_t p = 0; p < Bs; p++) {
    const uint32_t b0 = (uint32_t)(p * SC), sm = b0 + (uint32_t)(R * tot);
    for (int k = 0; k < R; k++) jobs.push_back(MsmJob{b0 + (uint32_t)k * tot, 0, tot});
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{sm + 2 * (uint32_t)k, 0, 2});                                  // (h, g_1)
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{sm + (uint32_t)(4 * m + k), (uint32_t)(n + 1), 1});            // enc_g
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{sm + (uint32_t)(6 * m + 2 * k), (uint32_t)(n + 2), 2});        // (ghat, pk)
  }
  {
    CK(cudaMemcpyAsync(d_g1_scal, h_scal.data(), Bs * (size_t)SC * 32, cudaMemcpyHostToDevice, st));
    CK(msm_run(ctx->ws, d_g1_scal, Bs * (size_t)SC, S->d_tab_ck, 1, jobs.data(), (int)jobs.size(), S->tab_c, d_g1_out, st, 0, -1, nb));
    launches += msm_last_launches(ctx->ws);
    const uint64_t totalE = Bs * 4 * (uint64_t)m;
    k_combine_E_batch<<<(unsigned)((totalE + 63) / 64), 64, 0, st>>>(d_ct_out, d_g1_out, JC, (uint32_t)(R + 2 * m), (uint32_t)(R + 4 * m),
                                                                     2 * m, totalE);
    CK(cudaGetLastError());
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, Bs * JC, st));
    CK(xyzz_to_canonical(d_ct_out, (uint32_t*)(d_canon + Bs * JC * kPointBytes), totalE, st));
    launches += 3;
    h_pts.resize((Bs * JC + totalE) * kPointBytes);
    CK(cudaMemcpyAsync(h_pts.data(), d_canon, h_pts.size(), cudaMemcpyDeviceToHost, st));
    CK(stream_wait(ctx, st));
    dev_done();
  }

  // ---- round D: Hadamard challenges; zero-argument rows, diagonals and commitments
  h_scal.assign(Bs * (size_t)SD * 8, 0);
  S->pool.run(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    const uint8_t* g1 = &h_pts[p * JC * kPointBytes];
    memcpy(proof + L.hB, g1, (size_t)m * kPointBytes);
    memcpy(proof + L.cb, g1 + kPointBytes * (size_t)(m - 1), kPointBytes);
    memcpy(proof + L.svpts, g1 + kPointBytes * (size_t)m, 3 * kPointBytes);
    memcpy(proof + L.mepts, g1 + kPointBytes * (size_t)(m + 3), (size_t)(2 * m + 1) * kPointBytes);
    memcpy(proof + L.meE, &h_pts[(Bs * JC + p * 4 * (size_t)m) * kPointBytes], 4 * (size_t)m * kPointBytes);
    h.fs.begin(); h.fs.feed_label("hadamard_argument"); h.fs.feed_points64(proof + L.cb, 1); h.fs.feed_points64(proof + L.hB, m); h.fs.end();
    h.xh = h.fs.challenge();
    h.yh = h.fs.challenge();
    h.xhp = h_powers(h.xh, m);
    // A' = (a0 | d_2..d_m | -1),  B' = (xh^i Bv_i (i = 1..m-1) | sum_{i=1}^{m-1} xh^i Bv_{i+1} | b_{m+1})
    h.Az.assign((size_t)(m + 1) * n, fr_zero());
    h.Bz.assign((size_t)(m + 1) * n, fr_zero());
    const fr minus1 = fr_neg(fr_one());
    for (int j = 0; j < n; j++) {
      h.Az[j] = h.z_a0[j];
      for (int k = 1; k < m; k++) h.Az[(size_t)k * n + j] = h.d[(size_t)k * n + j];
      h.Az[(size_t)m * n + j] = minus1;
      fr last = fr_zero();
      for (int i = 1; i < m; i++) {
        h.Bz[(size_t)(i - 1) * n + j] = fr_mul(h.xhp[i], h.Bv[(size_t)(i - 1) * n + j]);
        last = fr_add(last, fr_mul(h.xhp[i], h.Bv[(size_t)i * n + j]));
      }
      h.Bz[(size_t)(m - 1) * n + j] = last;
      h.Bz[(size_t)m * n + j] = h.z_bm1[j];
    }
    std::vector<fr> yp((size_t)n), dk((size_t)2 * m + 1, fr_zero());
    fr acc = fr_one();
    for (int j = 0; j < n; j++) { acc = fr_mul(acc, h.yh); yp[j] = acc; }
    for (int i = 0; i <= m; i++)
      for (int jj = 0; jj <= m; jj++) {
        fr v = fr_zero();
        for (int tt = 0; tt < n; tt++)
          v = fr_add(v, fr_mul(fr_mul(h.Az[(size_t)i * n + tt], h.Bz[(size_t)jj * n + tt]), yp[tt]));
        int k = i + m - jj;
        dk[k] = fr_add(dk[k], v);
      }
    uint32_t* base = &h_scal[p * (size_t)SD * 8];
    put_fr(base, h.z_r0);
    for (int j = 0; j < n; j++) put_fr(base + 8 * (size_t)(1 + j), h.z_a0[j]);
    put_fr(base + 8 * (size_t)tot, h.z_sm1);
    for (int j = 0; j < n; j++) put_fr(base + 8 * (size_t)(tot + 1 + j), h.z_bm1[j]);
    for (int k = 0; k <= 2 * m; k++) {
      put_fr(base + 8 * (size_t)(2 * tot + 2 * k), h.z_t[k]);
      put_fr(base + 8 * (size_t)(2 * tot + 2 * k + 1), dk[k]);
    }
  });
  jobs.clear();
  for (size_t p = 0; p < Bs; p++) {
    const uint32_t b0 = (uint32_t)(p * SD);
    jobs.push_back(MsmJob{b0, 0, tot});
    jobs.push_back(MsmJob{b0 + tot, 0, tot});
    for (int k = 0; k <= 2 * m; k++) jobs.push_back(MsmJob{b0 + 2 * tot + 2 * (uint32_t)k, 0, 2});
  }
  if ((rc = run_commit_jobs(Bs * (size_t)SD, jobs.size(), Bs * (size_t)JD)) != MP_OK) return rc;

  // ---- remaining challenges and all responses (host)
  S->pool.run(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    memcpy(proof + L.zpts, &h_pts[p * (size_t)JD * kPointBytes], (size_t)JD * kPointBytes);
    h.fs.begin(); h.fs.feed_label("zero_argument"); h.fs.feed_points64(proof + L.zpts, 2 * (size_t)m + 3); h.fs.end();
    const fr xz = h.fs.challenge();
    h.fs.begin(); h.fs.feed_label("single_value_product_argument"); h.fs.feed_points64(proof + L.svpts, 3); h.fs.end();
    const fr xs = h.fs.challenge();
    h.fs.begin(); h.fs.feed_label("multi_exponentiation_argument");
    h.fs.feed_points64(proof + L.mepts, 2 * (size_t)m + 1); h.fs.feed_points64(proof + L.meE, 4 * (size_t)m); h.fs.end();
    const fr xm = h.fs.challenge();
    const std::vector<fr> xzp = h_powers(xz, 2 * m + 1), xmp = h_powers(xm, 2 * m);
    for (int j = 0; j < n; j++) {
      fr za = fr_zero(), zb = fr_zero(), ma = h.me_a0[j];
      for (int i = 0; i <= m; i++) {
        za = fr_add(za, fr_mul(xzp[i], h.Az[(size_t)i * n + j]));
        zb = fr_add(zb, fr_mul(xzp[m - i], h.Bz[(size_t)i * n + j]));
      }
      for (int k = 1; k <= m; k++) ma = fr_add(ma, fr_mul(xmp[k], h.b[(size_t)(k - 1) * n + j]));
      h_fr_out(za, proof + L.za + 32 * (size_t)j);
      h_fr_out(zb, proof + L.zb + 32 * (size_t)j);
      h_fr_out(ma, proof + L.mea + 32 * (size_t)j);
      h_fr_out(fr_add(fr_mul(xs, h.col[j]), h.sv_d[j]), proof + L.sva + 32 * (size_t)j);
      h_fr_out(fr_add(fr_mul(xs, h.bk[j]), h.sv_delta[j]), proof + L.svb + 32 * (size_t)j);
    }
    std::vector<fr> rext((size_t)m + 1), sext((size_t)m + 1), xr((size_t)m + 1);
    rext[0] = h.z_r0;
    for (int i = 1; i < m; i++) rext[i] = h.t[i];
    rext[m] = fr_zero();
    for (int i = 1; i < m; i++) sext[i - 1] = fr_mul(h.xhp[i], h.sv[i - 1]);
    sext[m - 1] = h_dot(h.xhp.data() + 1, h.sv.data() + 1, m - 1);
    sext[m] = h.z_sm1;
    for (int j = 0; j <= m; j++) xr[j] = xzp[m - j];
    h_fr_out(h_dot(xzp.data(), rext.data(), m + 1), proof + L.zr);
    h_fr_out(h_dot(xr.data(), sext.data(), m + 1), proof + L.zs);
    h_fr_out(h_dot(xzp.data(), h.z_t.data(), 2 * m + 1), proof + L.zt);
    h_fr_out(fr_add(fr_mul(xs, h.s_prod), h.sv_rd), proof + L.svr);
    h_fr_out(fr_add(fr_mul(xs, h.sv_sx), h.sv_s1), proof + L.svs);
    std::vector<fr> mr((size_t)m + 1);
    mr[0] = h.me_r0;
    for (int j = 1; j <= m; j++) mr[j] = h.s[j - 1];
    h_fr_out(h_dot(xmp.data(), mr.data(), m + 1), proof + L.mer);
    h_fr_out(h_dot(xmp.data(), h.me_b.data(), 2 * m), proof + L.meb);
    h_fr_out(h_dot(xmp.data(), h.me_s.data(), 2 * m), proof + L.mes);
    h_fr_out(h_dot(xmp.data(), h.me_tau.data(), 2 * m), proof + L.metau);
  });
  int bad = 0;
  host_done();
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(stream_wait(ctx, st));
  dev_done();
  if (trace_on) fprintf(stderr, "prove_sub_batch: %zu proofs, %d threads: host phases %.2f ms, device phases %.2f ms\n", Bs, threads, t_host, t_dev);
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "the public key is not a canonical point of the curve");
  ctx->launches = launches;
  return MP_OK;
}


int32_t shuffle_prove_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                            const uint8_t* rhos, const uint8_t* rands, uint64_t B, uint8_t* out_decks,
                            uint8_t* proofs, int32_t host_threads, const void* d_decks) {
  if (!ctx || !pk || (B && (!decks || !perms || !rhos || !rands || !out_decks || !proofs))) return MP_ERR_INVALID_ARG;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n, plen = shuffle_proof_len(m, n), rlen = shuffle_randomness_len(m, n) * 32;
  int P = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
#ifdef MP_CURVE_BLS12_377
  if (N <= small_deck_max()) {
#else
  if (N <= small_deck_max() && !getenv("MP_BATCH_WORKERS")) {
#endif
    // small decks: lockstep over sub-batches (job grid <= 65535 per launch; bounded staging memory)
    const int threads = std::max(1, std::min(P, 64));
    const size_t per_proof_jobs = (size_t)(m + 4 + 6 * m);
    size_t sub = std::min<size_t>({(size_t)4096, (size_t)60000 / per_proof_jobs, std::max<size_t>(1, ((size_t)1 << 22) / N)});
    sub = std::max<size_t>(sub, 1);
    return run_chunks(ctx, B, sub, [&](mp_ctx* w, uint64_t p0, size_t Bs) {
      return prove_sub_batch(w, pk, decks + p0 * N * kCtBytes, perms + p0 * N, rhos + p0 * N * 32, rands + p0 * rlen, Bs,
                             out_decks + p0 * N * kCtBytes, proofs + p0 * plen, threads);
    });
  }
#ifdef MP_CURVE_BLS12_377
  // second curve: the lockstep prover above is the whole implementation (host-scalar form; the device-scalar
  // large-deck prover with its Karatsuba plan is built for the Stark curve only)
  return ctx->fail(MP_ERR_INVALID_ARG, "decks above %zu cards are not supported on this curve", small_deck_max());
#else
  // large decks (or MP_BATCH_WORKERS set): concurrent worker contexts running the single-proof path.
  // For 2^16-card decks a few workers are enough to hide each proof's serial Blake2s statement
  // absorb (host) behind the other proofs' kernels (device).
  // worker contexts for large decks: enough proofs in flight that the device always has kernels queued while other
  // proofs are in their host phases (statement hash, the four challenge round trips); each context owns ~2.3 GB of
  // MSM workspace at 2^16 cards.  Measured at 2^16 cards on one B200 (16 vCPUs): 30.2 / 34.2 / 36.5 / 36.4 proofs/s
  // with 3 / 6 / 8 / 12 contexts; host_threads caps it when several ranks share a host
  static const uint64_t large_workers = [] { const char* e = getenv("MP_PROVE_WORKERS"); return e ? strtoull(e, nullptr, 10) : 8ull; }();
  // (worker contexts sleep while their kernels run: for real large decks their number is not the caller's thread budget;
  //  the budget decides whether each worker hashes its own statement head or the heads are shared, StatementHashes)
  if (N > small_deck_max()) P = host_threads == 1 ? 1 : (int)std::max<uint64_t>(1, std::min<uint64_t>(large_workers, B));
  else P = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)P, (uint64_t)32, B}));
  StatementHashes hashes;
  const bool shared = share_statement_hashes(host_threads, P);
  if (shared) hashes.start(S, pk, decks, nullptr, N, B);
  return run_on_workers(ctx, P, B, [&](mp_ctx* w, uint64_t i) {
    const void* d_shuffled = nullptr;
    Transcript fs;
    int32_t st = shuffle_remask(w, pk, decks + i * N * 128, perms + i * N, rhos + i * N * 32, N, out_decks + i * N * 128,
                                d_decks ? (const uint8_t*)d_decks + i * N * 128 : nullptr, &d_shuffled, shared ? nullptr : &fs);
    int l = w->launches;
    if (st == MP_OK) {
      if (shared) fs.adopt_pending(hashes.wait(i));   // the head of the statement, hashed with the other decks'
      st = shuffle_prove(w, pk, decks + i * N * 128, out_decks + i * N * 128, perms + i * N, rhos + i * N * 32,
                         rands + i * rlen, proofs + i * plen, d_shuffled, &fs);
      w->launches += l;
    }
    return st;
  }, /*sleeping_waits=*/P > 1);
#endif
}

}  // namespace mp
