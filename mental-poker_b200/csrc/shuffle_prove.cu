// ShuffleArgument::prove for one deck with device-side scalar vectors (the 2^16-card path;
// reference call site mod.rs:409-415).  See shuffle.cuh for the design notes.
#include <chrono>

#include "shuffle_internal.cuh"

namespace mp {

// E_k = diag_k + Enc(b_k*ghat; tau_k):  E[2k] += c1[k] (= tau_k*g),  E[2k+1] += c2[k] (= b_k*ghat + tau_k*pk)
__global__ void __launch_bounds__(64) k_combine_E(xyzz* __restrict__ E, const xyzz* __restrict__ c1,
                                                  const xyzz* __restrict__ c2, int two_m) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 2 * two_m) return;
  xyzz x = E[g], y = (g & 1) ? c2[g >> 1] : c1[g >> 1];
  xyzz_add(x, y);
  E[g] = x;
}



// MP_TRACE=1: host-side timestamps (ms since entry) of the prover's synchronisation points, on stderr
struct ProveTrace {
  bool on = getenv("MP_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) const {
    nvtx_mark(what);
    if (on) fprintf(stderr, "[prove] %8.3f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
  }
};

int32_t shuffle_prove(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2, const uint32_t* perm,
                      const uint8_t* rho, const uint8_t* rand, uint8_t* proof_out, const void* deck2_src, Transcript* fs_started) {
  if (!ctx || !pk || !deck || !deck2 || !perm || !rho || !rand || !proof_out) return MP_ERR_INVALID_ARG;
  if (!deck2_src) deck2_src = deck2;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  NvtxRange nvtx("shuffle_prove");
  const ProveTrace trace;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  const Layout L(m, n);
  for (size_t i = 0; i < N; i++)
    if (perm[i] >= N) return ctx->fail(MP_ERR_INVALID_ARG, "permutation entry %zu out of range", i);
  RandCursor rc{rand};
  cudaStream_t st = ctx->stream;
  int32_t rcode;

  // ---- device buffers
  const size_t rows_max = (size_t)std::max(2 * m + 1, m + 4);
  const size_t T2 = N + 2;  // CT arena: deck2 | (g, pk) | (O, ghat)
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, T2 * 128);
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, T2 * 2 * sizeof(affine));
  uint32_t* d_perm = (uint32_t*)ctx->scratch(sPerm, N * 4);
  fr* d_rho = (fr*)ctx->scratch(sRho, N * 32 + 64);
  fr* d_a = (fr*)ctx->scratch(sFrA, N * sizeof(fr));
  fr* d_Ame = (fr*)ctx->scratch(sFrAme, (N + n) * sizeof(fr));        // rows: a0_me | b chunk 1..m
  fr* d_b = d_Ame + n;
  fr* d_Az = (fr*)ctx->scratch(sFrD, (N + n) * sizeof(fr));           // zero-arg rows: a0_z | d rows 1..m-1 | -1
  fr* d_d0 = (fr*)ctx->scratch(sFrTmp2, N * sizeof(fr));              // d (all m rows)
  fr* d_Bv = (fr*)ctx->scratch(sFrBv, N * sizeof(fr));
  fr* d_Bz = (fr*)ctx->scratch(sFrB, (N + n) * sizeof(fr));           // zero-arg rows: x^i Bv[i-1] | dlast | b_{m+1}
  fr* d_xpow = (fr*)ctx->scratch(sFrXpow, N * sizeof(fr));
  fr* d_pairs = (fr*)ctx->scratch(sFrPairs, (size_t)(m + 1) * (m + 1) * sizeof(fr) + (4 * (size_t)m + 8) * sizeof(fr));
  fr* d_partials = (fr*)ctx->scratch(sPartials, sizeof(fr) * (std::max(fr_powers_blocks(N), fr_reduce_blocks(N)) + 4));
  fr* d_rows = (fr*)ctx->scratch(sFrTmp0, (4 * (size_t)n + 64) * sizeof(fr));  // svp rows (3 x n) + response vectors
  uint32_t* d_g1_scal = (uint32_t*)ctx->scratch(sG1Scal, (rows_max * (n + 1) + 12 * (size_t)m + 64) * 32);
  xyzz* d_g1_out = (xyzz*)ctx->scratch(sG1Out, (8 * (size_t)m + 16) * sizeof(xyzz));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, (N + n + 4 * (size_t)m + 8) * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, 8 * (size_t)m * sizeof(xyzz));
  xyzz* d_enc = d_ct_out + 4 * (size_t)m;  // Enc(b_k*ghat; tau_k) parts: 2m c1 values, then 2m c2 values
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sCanonOut, (8 * (size_t)m + 16) * 64);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_perm); NEED(d_rho); NEED(d_a); NEED(d_Ame); NEED(d_Az); NEED(d_d0);
  NEED(d_Bv); NEED(d_Bz); NEED(d_xpow); NEED(d_pairs); NEED(d_partials); NEED(d_rows); NEED(d_g1_scal); NEED(d_g1_out);
  NEED(d_ct_scal); NEED(d_ct_out); NEED(d_canon); NEED(d_bad);
  const size_t pin_bytes = (8 * (size_t)m + 16) * 64 + (4 * (size_t)n + 4 * (size_t)m + 64) * 32;
  uint8_t* h_pin = pinned(S, pin_bytes);
  if (!h_pin) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");

  // ---- uploads that do not depend on any challenge
  CK(cudaStreamWaitEvent(st, S->ev_bulk_done, 0));  // (bulk work of a call that failed midway, if any)
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  if (deck2_src != (const void*)d_ct_canon)  // (shuffle_and_remask leaves the remasked deck right here)
    CK(cudaMemcpyAsync(d_ct_canon, deck2_src, N * 128, cudaMemcpyDefault, st));
  {
    uint8_t tail[256];
    memset(tail, 0, sizeof tail);
    memcpy(tail, S->enc_g, 64);
    memcpy(tail + 64, pk, 64);
    memcpy(tail + 192, S->ghat, 64);
    CK(cudaMemcpyAsync(d_ct_canon + N * 128, tail, 256, cudaMemcpyHostToDevice, st));
  }
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, T2 * 2, d_bad, st));
  if (!S->ck_pk_valid || memcmp(S->ck_pk, pk, 64) != 0) {  // pk column of the fixed-base table (cached)
    CK(cudaMemcpyAsync(S->d_ck + (n + 3), d_ct_mont + 2 * N + 1, sizeof(affine), cudaMemcpyDeviceToDevice, st));
    CK(msm_build_table(ctx->ws, S->d_ck, (uint32_t)(n + 4), (uint32_t)(n + 3), 1, S->tab_c, S->d_tab_ck, st));
    memcpy(S->ck_pk, pk, 64);
    S->ck_pk_valid = true;
    ctx->launches += 2;
  }
  CK(cudaMemcpyAsync(d_perm, perm, N * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_xpow, rho, N * 32, cudaMemcpyHostToDevice, st));  // staging: canonical rho
  CK(fr_from_canonical_vec((const uint32_t*)d_xpow, d_rho, N, st));
  ctx->launches += 2;

  // ---- round A: c_A[k] = com(chunk_k(a); r_k),  a_i = perm[i] + 1
  const std::vector<fr> r = rc.vec(m), s = rc.vec(m);
  // all remaining prover randomness, drawn now in the B.6 order (the draws depend on no challenge)
  const fr s_prod = rc.one();                                   // B.2: blinding of c_b
  std::vector<fr> sv((size_t)m);                                // B.3: s_1 = t_1, s_m = s_prod, rest random
  // sv[0] = t[0] is filled in once y is known (round C)
  for (int i = 1; i < m - 1; i++) sv[i] = rc.one();
  sv[m - 1] = s_prod;
  // B.4 randomness (drawn now to respect the B.6 order; used in round D)
  const std::vector<fr> z_a0 = rc.vec(n), z_bm1 = rc.vec(n);
  const fr z_r0 = rc.one(), z_sm1 = rc.one();
  std::vector<fr> z_t((size_t)2 * m + 1);
  for (int k = 0; k <= 2 * m; k++) z_t[k] = (k != m + 1) ? rc.one() : fr_zero();
  // B.5 randomness
  const std::vector<fr> sv_d = rc.vec(n);
  const fr sv_rd = rc.one();
  std::vector<fr> sv_delta((size_t)n);
  sv_delta[0] = sv_d[0];
  for (int i = 1; i < n - 1; i++) sv_delta[i] = rc.one();
  sv_delta[n - 1] = fr_zero();
  const fr sv_s1 = rc.one(), sv_sx = rc.one();
  // B.5' randomness
  const std::vector<fr> me_a0 = rc.vec(n);
  const fr me_r0 = rc.one();
  std::vector<fr> me_b((size_t)2 * m), me_s((size_t)2 * m), me_tau((size_t)2 * m);
  for (int k = 0; k < 2 * m; k++) {
    if (k == m) { me_b[k] = fr_zero(); me_s[k] = fr_zero(); me_tau[k] = fr_zero(); /* tau_m = rho*, set below */ }
    else { me_b[k] = rc.one(); me_s[k] = rc.one(); me_tau[k] = rc.one(); }
  }
  if (rc.i != shuffle_randomness_len(m, n)) return ctx->fail(MP_ERR_INVALID_ARG, "internal: randomness count mismatch");

  fr* d_blind = d_pairs;  // small scratch for blinding factors (<= 4m + 8 elements at the end of d_pairs)
  d_blind = d_pairs + (size_t)(m + 1) * (m + 1);
  CK(fr_perm_vectors(d_perm, nullptr, N, d_a, nullptr, st));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(d_blind, r.data(), sizeof(fr) * m, cudaMemcpyHostToDevice, st));
  if ((rcode = commit_rows_device(ctx, d_a, n, d_blind, m, n, d_g1_scal, d_g1_out)) != MP_OK) return rcode;
  CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, m, st));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(proof_out + L.cA, d_canon, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
  CK(mark_record(ctx, S->ev, st));
  // Every shuffled-deck point is multiplied by m + 1 scalar rows in the diagonal MSMs, so
  // pre-shifting it once (table[w] = 2^(c w) * point) pays: all windows of a job then share ONE
  // bucket set -- one bucket reduction per job instead of W, no fold doublings, and a wider
  // window (fewer entries).  The table depends on no challenge: it is queued behind c_A and
  // runs on the GPU while the host hashes the statement.
  const int c_diag = msm_pick_table_window(N / 2 + 1);
  // From m ~ 16 on, Karatsuba on the row index (diag.cu) needs fewer bucket additions than the
  // m(m+1) row products of the schoolbook form; its leaf point rows depend on no challenge either.
  const bool kara = diag_karatsuba_selected(m, n, c_diag);
  affine* d_ct_tab = nullptr;
  if (kara) {
    if ((rcode = diag_karatsuba_points(ctx, d_ct_mont, st)) != MP_OK) return rcode;
  } else {
    d_ct_tab = (affine*)ctx->scratch(sCtTable, (size_t)msm_num_windows(c_diag) * T2 * 2 * sizeof(affine));
    NEED(d_ct_tab);
    CK(msm_build_table(ctx->ws, d_ct_mont, (uint32_t)(T2 * 2), 0, (uint32_t)(N * 2), c_diag, d_ct_tab, st));
    ctx->launches += 2;
  }
  // The statement hash (17 MB through one Blake2s chain at 2^16 cards) is the longest serial item of
  // the whole proof: everything but c_A is hashed before the host waits for the GPU.
  Transcript fs_local;
  Transcript& fs = fs_started ? *fs_started : fs_local;
  if (!fs_started) absorb_statement_head(fs, S, pk, deck, N);
  absorb_statement_deck2(fs, deck2, N);
  CK(mark_wait(ctx, S->ev));  // c_A is on the host; the table / leaf-row kernels continue
  trace.mark("statement hashed up to c_A; c_A on host");
  absorb_statement_tail(fs, proof_out + L.cA, m);
  const fr x = fs.challenge();

  // ---- round B: b_i = x^{perm[i]+1}, c_B[k] = com(chunk_k(b); s_k)
  CK(fr_powers(h_pow2_table(x), N, nullptr, d_xpow, nullptr, d_partials, nullptr, st));
  CK(fr_perm_vectors(d_perm, d_xpow, N, nullptr, d_b, st));
  ctx->launches += 2;

  // ---- the 2m diagonal ciphertext products E_k (K2; multi-exponentiation first message, B.5') are the
  // prover's dominant device cost.  They need x and nothing else, and their result only enters the
  // LAST transcript absorb -- so they go to the bulk stream now; the small launches and host round
  // trips of rounds B, C and D (main stream) fill in around them.
  CK(cudaMemcpyAsync(d_Ame, me_a0.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
  CK(fr_to_canonical_vec(d_Ame, d_ct_scal, N + n, st));  // scalar arena: rows a0 | b_1..b_m
  ctx->launches += 1;
  CK(cudaEventRecord(S->ev_bulk_go, st));
  CK(cudaStreamWaitEvent(S->bulk, S->ev_bulk_go, 0));
  trace.mark("x known; b rows queued");
  if (kara) {
    if ((rcode = diag_karatsuba_products(ctx, d_ct_scal, d_ct_out, S->bulk, S->bulk_ws)) != MP_OK) return rcode;
  } else {
    std::vector<MsmJob> diag((size_t)2 * m);
    for (int k = 0; k < 2 * m; k++) {
      int i0 = std::max(1, m - k), i1 = std::min(m, 2 * m - k);
      diag[k] = MsmJob{(uint32_t)((size_t)(k - m + i0) * n), (uint32_t)((size_t)(i0 - 1) * n), (uint32_t)((size_t)(i1 - i0 + 1) * n)};
    }
    // one launch sequence normally; very large decks are split so that a call stays below the
    // 2^32-entry limit of the sort (entries = terms * windows)
    const uint64_t max_terms = ((1ull << 31) / (uint64_t)msm_num_windows(c_diag));
    for (int k0 = 0; k0 < 2 * m;) {
      int k1 = k0;
      uint64_t terms = 0;
      while (k1 < 2 * m && (k1 == k0 || terms + diag[k1].len <= max_terms)) terms += diag[k1++].len;
      CK(msm_run(S->bulk_ws, d_ct_scal, N + n, d_ct_tab, 2, diag.data() + k0, k1 - k0, c_diag, d_ct_out + 2 * (size_t)k0, S->bulk, 0, -1,
                 (uint32_t)T2));
      ctx->launches += msm_last_launches(S->bulk_ws);
      k0 = k1;
    }
  }
  CK(cudaEventRecord(S->ev_bulk_done, S->bulk));
  trace.mark("diagonal products queued (bulk stream)");

  CK(cudaMemcpyAsync(d_blind, s.data(), sizeof(fr) * m, cudaMemcpyHostToDevice, st));
  if ((rcode = commit_rows_device(ctx, d_b, n, d_blind, m, n, d_g1_scal, d_g1_out)) != MP_OK) return rcode;
  CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, m, st));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(proof_out + L.cB, d_canon, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
  CK(stream_wait(ctx, st));
  trace.mark("round B: c_B on host");
  fs.begin(); fs.feed_label("shuffle_argument_b"); fs.feed_points64(proof_out + L.cB, m); fs.end();
  const fr y = fs.challenge();
  const fr z = fs.challenge();

  // ---- round C: everything whose commitments depend only on (x, y, z)
  // C.1  d = y*a + b - z;  column prefix products Bv;  rho* = -sum rho_i b_i
  std::vector<fr> t((size_t)m);
  for (int k = 0; k < m; k++) t[k] = fr_add(fr_mul(y, r[k]), s[k]);
  sv[0] = t[0];
  {
    fr yz[2] = {y, z};
    CK(cudaMemcpyAsync(d_blind, yz, sizeof yz, cudaMemcpyHostToDevice, st));
    CK(fr_affine_comb(d_a, d_b, d_blind, N, d_d0, st));
    CK(fr_column_prefix_products(d_d0, m, n, d_Bv, st));
    CK(fr_dot(d_rho, d_b, N, d_partials, d_partials + fr_reduce_blocks(N), st));
    ctx->launches += 4;
  }
  // bring the last product row (the SVP witness) and rho* to the host
  fr* h_col = reinterpret_cast<fr*>(h_pin);
  CK(cudaMemcpyAsync(h_col, d_Bv + (size_t)(m - 1) * n, sizeof(fr) * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_col + n, d_partials + fr_reduce_blocks(N), sizeof(fr), cudaMemcpyDeviceToHost, st));
  CK(mark_record(ctx, S->ev, st));

  CK(mark_wait(ctx, S->ev));  // col / rho* are on the host
  trace.mark("round C.1: product column on host");
  std::vector<fr> col(h_col, h_col + n);
  const fr rho_star = fr_neg(h_col[n]);
  me_tau[m] = rho_star;

  // C.4  SVP first message on the host side of the scalars (O(n)), committed on the device.
  cudaStream_t sa = st;
  std::vector<fr> bk((size_t)n);
  bk[0] = col[0];
  for (int i = 1; i < n; i++) bk[i] = fr_mul(bk[i - 1], col[i]);
  {
    std::vector<fr> rows3((size_t)3 * n, fr_zero());
    for (int i = 0; i < n; i++) rows3[i] = sv_d[i];
    for (int i = 0; i + 1 < n; i++) {
      rows3[(size_t)n + i] = fr_neg(fr_mul(sv_delta[i], sv_d[i + 1]));
      rows3[(size_t)2 * n + i] = fr_sub(fr_sub(sv_delta[i + 1], fr_mul(col[i + 1], sv_delta[i])), fr_mul(bk[i], sv_d[i + 1]));
    }
    CK(cudaMemcpyAsync(d_rows, rows3.data(), sizeof(fr) * 3 * n, cudaMemcpyHostToDevice, sa));
  }
  // C.5  ONE G1 batch (one Pippenger launch sequence = one fold latency):
  //      rows      Hadamard c_B[0..m) = com(Bv[i]; sv[i]) (c_B[0] = c_D[0], c_B[m-1] = c_b), SVP c_d,
  //                c_delta, c_Delta, multi-exp c_A0                              (n+1 terms each)
  //      pairs     multi-exp c_B_k = s_k*h + b_k*g_1                               (2 terms)
  //      enc c1/c2 Enc(b_k*ghat; tau_k) = (tau_k*g, b_k*ghat + tau_k*pk)           (1 / 2 terms)
  //      (kept in d_enc: E_k = diag_k + enc_k is formed once the bulk stream has delivered diag_k)
  {
    const int R = m + 4;
    std::vector<fr> blinds((size_t)R);
    for (int i = 0; i < m; i++) blinds[i] = sv[i];
    blinds[m] = sv_rd; blinds[m + 1] = sv_s1; blinds[m + 2] = sv_sx; blinds[m + 3] = me_r0;
    CK(cudaMemcpyAsync(d_blind, blinds.data(), sizeof(fr) * R, cudaMemcpyHostToDevice, sa));
    const uint64_t tot = (uint64_t)(n + 1);
    CK(commit_scalars_launch(d_Bv, n, d_blind, m, n, n, d_g1_scal, sa));
    CK(commit_scalars_launch(d_rows, n, d_blind + m, 3, n, n, d_g1_scal + tot * m * 8, sa));
    CK(commit_scalars_launch(d_Ame, n, d_blind + m + 3, 1, n, n, d_g1_scal + tot * (m + 3) * 8, sa));
    ctx->launches += 3;
    const size_t nsmall = 10 * (size_t)m;  // 2m pairs + 2m singles + 2m pairs
    std::vector<uint32_t> sc(nsmall * 8);
    for (int k = 0; k < 2 * m; k++) {
      fr_to_canonical(me_s[k], &sc[(size_t)(2 * k) * 8]);
      fr_to_canonical(me_b[k], &sc[(size_t)(2 * k + 1) * 8]);
      fr_to_canonical(me_tau[k], &sc[(size_t)(4 * m + k) * 8]);
      fr_to_canonical(me_b[k], &sc[(size_t)(6 * m + 2 * k) * 8]);
      fr_to_canonical(me_tau[k], &sc[(size_t)(6 * m + 2 * k + 1) * 8]);
    }
    const uint32_t base = (uint32_t)(tot * R);
    CK(cudaMemcpyAsync(d_g1_scal + (size_t)base * 8, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice, sa));
    std::vector<MsmJob> jobs;
    for (int k = 0; k < R; k++) jobs.push_back(MsmJob{(uint32_t)(k * tot), 0, (uint32_t)tot});
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{base + 2 * k, 0, 2});                               // (h, g_1)
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{base + 4 * m + k, (uint32_t)(n + 1), 1});           // enc_g
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{base + 6 * m + 2 * k, (uint32_t)(n + 2), 2});       // (ghat, pk)
    CK(msm_run(ctx->ws, d_g1_scal, base + nsmall, S->d_tab_ck, 1, jobs.data(), (int)jobs.size(), S->tab_c, d_g1_out, sa, 0, -1,
               (uint32_t)(n + 4)));
    ctx->launches += msm_last_launches(ctx->ws);
    CK(cudaMemcpyAsync(d_enc, d_g1_out + R + 2 * m, 4 * (size_t)m * sizeof(xyzz), cudaMemcpyDeviceToDevice, st));
    uint8_t* d_canon2 = d_canon + 4 * (size_t)m * 64;
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon2, (size_t)R + 2 * m, st));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(proof_out + L.hB, d_canon2, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.svpts, d_canon2 + (size_t)m * 64, 3 * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.mepts, d_canon2 + (size_t)(m + 3) * 64, (size_t)(2 * m + 1) * 64, cudaMemcpyDeviceToHost, st));
  }
  CK(stream_wait(ctx, st));
  trace.mark("round C.5: commitment batch on host");
  memcpy(proof_out + L.cb, proof_out + L.hB + 64 * (size_t)(m - 1), 64);  // c_b = c_B[m-1]
  fs.begin(); fs.feed_label("hadamard_argument"); fs.feed_points64(proof_out + L.cb, 1); fs.feed_points64(proof_out + L.hB, m); fs.end();
  const fr xh = fs.challenge();
  const fr yh = fs.challenge();

  // ---- round D: zero argument (B.4) on A' = (a0 | d_2..d_m | -1), B' = (xh^i Bv_i | dlast | b_{m+1})
  const std::vector<fr> xhp = h_powers(xh, m);
  {
    // A rows
    CK(cudaMemcpyAsync(d_Az, z_a0.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_Az + n, d_d0 + n, sizeof(fr) * (N - n), cudaMemcpyDeviceToDevice, st));
    std::vector<fr> m1((size_t)n, fr_neg(fr_one()));
    CK(cudaMemcpyAsync(d_Az + N, m1.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
    // B rows: row i-1 = xh^i * Bv[i-1] (i = 1..m-1); row m-1 = sum_{i=1}^{m-1} xh^i Bv[i]; row m = b_{m+1}
    SmallUpload up;
    std::vector<fr> coef(xhp.begin() + 1, xhp.end());  // xh^1 .. xh^{m-1}
    size_t o_coef = up.add_frs(coef);
    std::vector<fr> yp((size_t)n);
    fr acc = fr_one();
    for (int j = 0; j < n; j++) { acc = fr_mul(acc, yh); yp[j] = acc; }
    size_t o_yp = up.add_frs(yp);
    uint8_t* d_small = (uint8_t*)ctx->scratch(sSmallUp, up.bytes.size() + 64);
    NEED(d_small);
    CK(cudaMemcpyAsync(d_small, up.bytes.data(), up.bytes.size(), cudaMemcpyHostToDevice, st));
    CK(fr_scale_rows(d_Bv, (const fr*)(d_small + o_coef), m - 1, n, d_Bz, st));
    CK(fr_lincomb_rows(d_Bv + n, n, (const fr*)(d_small + o_coef), m - 1, n, d_Bz + (size_t)(m - 1) * n, st));
    CK(cudaMemcpyAsync(d_Bz + N, z_bm1.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
    fr* d_dk = d_pairs + (size_t)(m + 1) * (m + 1) + 2 * (size_t)m + 4;  // 2m+1 diagonal sums
    CK(fr_bilinear_diagonals(d_Az, d_Bz, (const fr*)(d_small + o_yp), m + 1, n, d_pairs, d_dk, st));
    ctx->launches += 4;
    // commitments: c_A0 = com(a0; r0), c_Bm1 = com(b_{m+1}; s_{m+1}), c_D_k = com(d_k; t_k)
    fr bl[2] = {z_r0, z_sm1};
    CK(cudaMemcpyAsync(d_blind, bl, sizeof bl, cudaMemcpyHostToDevice, st));
    uint64_t tot = (uint64_t)(n + 1);
    CK(commit_scalars_launch(d_Az, n, d_blind, 1, n, n, d_g1_scal, st));
    CK(commit_scalars_launch(d_Bz + N, n, d_blind + 1, 1, n, n, d_g1_scal + tot * 8, st));
    CK(cudaGetLastError());
    uint32_t* d_pairs_scal = d_g1_scal + 2 * tot * 8;  // (t_k, d_k) pairs
    std::vector<uint32_t> tk((size_t)(2 * m + 1) * 8);
    for (int k = 0; k <= 2 * m; k++) fr_to_canonical(z_t[k], &tk[(size_t)k * 8]);
    // t_k at even slots (strided copy), d_k at odd slots
    CK(cudaMemcpy2DAsync(d_pairs_scal, 64, tk.data(), 32, 32, 2 * (size_t)m + 1, cudaMemcpyHostToDevice, st));
    CK(fr_scatter_canonical(d_dk, 2 * (size_t)m + 1, d_pairs_scal, 1, 2, st));
    ctx->launches += 3;
    std::vector<MsmJob> jobs = {MsmJob{0, 0, (uint32_t)tot}, MsmJob{(uint32_t)tot, 0, (uint32_t)tot}};
    for (int k = 0; k <= 2 * m; k++) jobs.push_back(MsmJob{(uint32_t)(2 * tot + 2 * k), 0, 2});
    CK(msm_run(ctx->ws, d_g1_scal, 2 * tot + 2 * (2 * (size_t)m + 1), S->d_tab_ck, 1, jobs.data(), (int)jobs.size(),
               S->tab_c, d_g1_out, st, 0, -1, (uint32_t)(n + 4)));
    ctx->launches += msm_last_launches(ctx->ws);
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, 2 * (size_t)m + 3, st));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(proof_out + L.zpts, d_canon, (2 * (size_t)m + 3) * 64, cudaMemcpyDeviceToHost, st));
    if (trace.on) { CK(stream_wait(ctx, st)); trace.mark("round D: zero-argument commitments on host"); }
    // join the bulk stream: E_k = diag_k + Enc(b_k*ghat; tau_k), needed for the last absorb below
    CK(cudaStreamWaitEvent(st, S->ev_bulk_done, 0));
    k_combine_E<<<(4 * m + 63) / 64, 64, 0, st>>>(d_ct_out, d_enc, d_enc + 2 * m, 2 * m);
    CK(cudaGetLastError());
    uint8_t* d_canonE = d_canon + (2 * (size_t)m + 4) * 64;
    CK(xyzz_to_canonical(d_ct_out, (uint32_t*)d_canonE, 4 * (size_t)m, st));
    ctx->launches += 2;
    CK(cudaMemcpyAsync(proof_out + L.meE, d_canonE, 4 * (size_t)m * 64, cudaMemcpyDeviceToHost, st));
    CK(stream_wait(ctx, st));
  }
  trace.mark("round D + E_k on host");
  fs.begin(); fs.feed_label("zero_argument"); fs.feed_points64(proof_out + L.zpts, 2 * (size_t)m + 3); fs.end();
  const fr xz = fs.challenge();
  fs.begin(); fs.feed_label("single_value_product_argument"); fs.feed_points64(proof_out + L.svpts, 3); fs.end();
  const fr xs = fs.challenge();
  fs.begin(); fs.feed_label("multi_exponentiation_argument");
  fs.feed_points64(proof_out + L.mepts, 2 * (size_t)m + 1); fs.feed_points64(proof_out + L.meE, 4 * (size_t)m); fs.end();
  const fr xm = fs.challenge();

  // ---- responses.  Device: the three O(N) row combinations; host: the O(m + n) rest.
  const std::vector<fr> xzp = h_powers(xz, 2 * m + 1);
  const std::vector<fr> xmp = h_powers(xm, 2 * m);
  {
    SmallUpload up;
    std::vector<fr> ca(xzp.begin(), xzp.begin() + m + 1);       // a = sum_{i=0}^{m} xz^i A'_i
    std::vector<fr> cb((size_t)m + 1);                          // b = sum_{j=0}^{m} xz^{m-j} B'_j
    for (int j = 0; j <= m; j++) cb[j] = xzp[m - j];
    std::vector<fr> cm(xmp.begin(), xmp.begin() + m + 1);       // a_me = sum_{j=0}^{m} xm^j Ame_j
    size_t o_a = up.add_frs(ca), o_b = up.add_frs(cb), o_m = up.add_frs(cm);
    uint8_t* d_small = (uint8_t*)ctx->scratch(sSmallUp, up.bytes.size() + 64);
    NEED(d_small);
    CK(cudaMemcpyAsync(d_small, up.bytes.data(), up.bytes.size(), cudaMemcpyHostToDevice, st));
    fr* d_resp = d_rows;  // 3 x n
    CK(fr_lincomb_rows(d_Az, n, (const fr*)(d_small + o_a), m + 1, n, d_resp, st));
    CK(fr_lincomb_rows(d_Bz, n, (const fr*)(d_small + o_b), m + 1, n, d_resp + n, st));
    CK(fr_lincomb_rows(d_Ame, n, (const fr*)(d_small + o_m), m + 1, n, d_resp + 2 * (size_t)n, st));
    uint32_t* d_resp_canon = d_g1_scal;
    CK(fr_to_canonical_vec(d_resp, d_resp_canon, 3 * (size_t)n, st));
    ctx->launches += 4;
    CK(cudaMemcpyAsync(proof_out + L.za, d_resp_canon, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.zb, d_resp_canon + (size_t)n * 8, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.mea, d_resp_canon + 2 * (size_t)n * 8, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  }
  // zero-argument blinding responses: r' = (r0, t_2..t_m, 0), s' = (xh^i sv_i.., sum xh^i sv_{i+1}, s_{m+1})
  {
    std::vector<fr> rext((size_t)m + 1), sext((size_t)m + 1), xr((size_t)m + 1);
    rext[0] = z_r0;
    for (int i = 1; i < m; i++) rext[i] = t[i];
    rext[m] = fr_zero();
    for (int i = 1; i < m; i++) sext[i - 1] = fr_mul(xhp[i], sv[i - 1]);
    sext[m - 1] = h_dot(xhp.data() + 1, sv.data() + 1, m - 1);
    sext[m] = z_sm1;
    for (int j = 0; j <= m; j++) xr[j] = xzp[m - j];
    h_fr_out(h_dot(xzp.data(), rext.data(), m + 1), proof_out + L.zr);
    h_fr_out(h_dot(xr.data(), sext.data(), m + 1), proof_out + L.zs);
    h_fr_out(h_dot(xzp.data(), z_t.data(), 2 * m + 1), proof_out + L.zt);
  }
  // SVP responses
  for (int i = 0; i < n; i++) {
    h_fr_out(fr_add(fr_mul(xs, col[i]), sv_d[i]), proof_out + L.sva + 32 * (size_t)i);
    h_fr_out(fr_add(fr_mul(xs, bk[i]), sv_delta[i]), proof_out + L.svb + 32 * (size_t)i);
  }
  h_fr_out(fr_add(fr_mul(xs, s_prod), sv_rd), proof_out + L.svr);
  h_fr_out(fr_add(fr_mul(xs, sv_sx), sv_s1), proof_out + L.svs);
  // multi-exp responses
  {
    std::vector<fr> rext((size_t)m + 1);
    rext[0] = me_r0;
    for (int j = 1; j <= m; j++) rext[j] = s[j - 1];
    h_fr_out(h_dot(xmp.data(), rext.data(), m + 1), proof_out + L.mer);
    h_fr_out(h_dot(xmp.data(), me_b.data(), 2 * m), proof_out + L.meb);
    h_fr_out(h_dot(xmp.data(), me_s.data(), 2 * m), proof_out + L.mes);
    h_fr_out(h_dot(xmp.data(), me_tau.data(), 2 * m), proof_out + L.metau);
  }
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(stream_wait(ctx, st));
  trace.mark("responses on host");
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck point or the public key is not a canonical point of the Stark curve");
  return MP_OK;
}


}  // namespace mp
