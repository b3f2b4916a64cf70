// verify_shuffle on the GPU (reference DLCards::verify_shuffle, mod.rs:420-443): single proof with
// device-side O(N) scalars, and the lockstep batch.  The checks themselves are built by the
// host-only shuffle_host.hpp.  See shuffle.cuh for the design notes.
#include "comm.cuh"
#include "shuffle_internal.cuh"

namespace mp {

// ------------------------------------------------------------------------------------------
// verify
// ------------------------------------------------------------------------------------------
int32_t shuffle_verify(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                       const uint8_t* proof, const void* deck_src, const void* deck2_src, StatementHashes* hashes,
                       uint64_t hash_index) {
  NvtxRange nvtx("shuffle_verify");
  if (!ctx || !pk || !deck || !deck2 || !proof) return MP_ERR_INVALID_ARG;
  if (!deck_src) deck_src = deck;     // host copy doubles as the transfer source
  if (!deck2_src) deck2_src = deck2;  // (a device pointer here means the deck is already resident in HBM)
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  const Layout L(m, n);
  if (!proof_scalars_canonical(proof, L))
    return ctx->fail(MP_ERR_NOT_CANONICAL, "a scalar of the proof is not below the group order");

  // ---- 1. start moving the decks (independent of every challenge)
  const size_t T = 2 * N + 2 * (size_t)m + 3;  // CT arena: deck | E_m | deck2 | E_0..E_{2m-1} | (g,pk) | (O,ghat)
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, T * 128);
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, T * 2 * sizeof(affine));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, T * 32);
  // one large proof across GPUs (mp_shuffle_verify_multi): the two ciphertext equations go to ranks 0 and 1, the
  // 2 x 128-byte results are all-gathered; the G1 jobs and the transcript run on every rank
  const int G = comm_collective(ctx) ? comm_size(ctx) : 1, rank = G > 1 ? comm_rank(ctx) : 0;
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, (size_t)std::max(4, 2 * G) * sizeof(xyzz));
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_ct_scal); NEED(d_ct_out); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  CK(cudaEventRecord(S->ev_fork, ctx->stream));  // the auxiliary stream may touch d_bad after this point
  CK(cudaMemcpyAsync(d_ct_canon, deck_src, N * 128, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_canon + N * 128, proof + L.meE + 128 * (size_t)m, 128, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_canon + (N + 1) * 128, deck2_src, N * 128, cudaMemcpyDefault, ctx->stream));
  {
    std::vector<uint8_t> tail((2 * (size_t)m + 2) * 128, 0);
    memcpy(tail.data(), proof + L.meE, 2 * (size_t)m * 128);
    uint8_t* q = tail.data() + 2 * (size_t)m * 128;
    memcpy(q, S->enc_g, 64);
    memcpy(q + 64, pk, 64);
    memcpy(q + 192, S->ghat, 64);  // (identity, ghat)
    CK(cudaMemcpyAsync(d_ct_canon + (2 * N + 1) * 128, tail.data(), tail.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, T * 2, d_bad, ctx->stream));
  ctx->launches += 1;

  // ---- 2. transcript: every challenge derives from statement + proof bytes
  const Challenges ch = derive_challenges(S, pk, deck, deck2, N, proof, L, hashes ? &hashes->wait(hash_index) : nullptr);
  const fr &x = ch.x, &y = ch.y, &z = ch.z, &xm = ch.xm;

  // ---- 3. the commitment-space checks as small G1 jobs (host builds O(m + n) scalars), issued on
  //         the auxiliary stream so that they overlap the ciphertext MSMs below
  TermList tl;
  HostChecks hc;
  append_g1_checks(tl, S, proof, L, ch, &hc);
  const int J = (int)tl.jobs.size();  // 8
  xyzz* d_g1_out = nullptr;
  CK(cudaStreamWaitEvent(S->aux, S->ev_fork, 0));
  int32_t st = run_g1_jobs(ctx, tl, &d_g1_out, d_bad, S->aux, S->aux_ws);
  if (st != MP_OK) return st;
  CK(cudaEventRecord(S->ev_join, S->aux));

  // ---- 4. O(N) scalar vectors on the device
  const std::vector<fr> me_a = h_frs(proof + L.mea, n);
  const fr me_b = h_fr(proof + L.meb), me_tau = h_fr(proof + L.metau);
  const std::vector<fr> xmp = h_powers(xm, 2 * m);
  SmallUpload up;
  fr yz[2] = {y, z};
  size_t o_yz = up.add(yz, sizeof yz);
  std::vector<fr> coef((size_t)m);
  for (int i = 1; i <= m; i++) coef[i - 1] = fr_neg(xmp[m - i]);
  size_t o_coef = up.add_frs(coef);
  size_t o_mea = up.add_frs(me_a);
  std::vector<uint32_t> tailsc((2 * (size_t)m + 2) * 8);
  for (int k = 0; k < 2 * m; k++) fr_to_canonical(xmp[k], &tailsc[8 * (size_t)k]);
  fr_to_canonical(fr_neg(me_tau), &tailsc[8 * (size_t)(2 * m)]);
  fr_to_canonical(fr_neg(me_b), &tailsc[8 * (size_t)(2 * m + 1)]);
  uint32_t minus_one[8];
  fr_to_canonical(fr_neg(fr_one()), minus_one);
  uint8_t* d_small = (uint8_t*)ctx->scratch(sSmallUp, up.bytes.size() + 64);
  fr* d_partials = (fr*)ctx->scratch(sPartials, sizeof(fr) * (fr_powers_blocks(N) + 2));
  NEED(d_small); NEED(d_partials);
  CK(cudaMemcpyAsync(d_small, up.bytes.data(), up.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_scal + N * 8, minus_one, 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_scal + (2 * N + 1) * 8, tailsc.data(), tailsc.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  fr* d_bstar = d_partials + fr_powers_blocks(N);
  CK(fr_powers(h_pow2_table(x), N, d_ct_scal, nullptr, (const fr*)(d_small + o_yz), d_partials, d_bstar, ctx->stream));
  CK(fr_outer_canonical((const fr*)(d_small + o_coef), (const fr*)(d_small + o_mea), m, n, d_ct_scal + (N + 1) * 8, ctx->stream));
  ctx->launches += 3;

  // ---- 5. the two ciphertext checks (K1): 4 N-term G1 MSMs in one batched launch sequence
  //   job 0:  sum x^i C_i - E_m                                        == O   (Chat == E_m)
  //   job 1:  sum x^k E_k - Enc(b*ghat; tau) - sum (x^{m-i} a_j) C'_ij  == O
  MsmJob ct_jobs[2] = {{0, 0, (uint32_t)(N + 1)}, {(uint32_t)(N + 1), (uint32_t)(N + 1), (uint32_t)(N + 2 * m + 2)}};
  if (G == 1) {
    CK(msm_run(ctx->ws, d_ct_scal, T, d_ct_mont, 2, ct_jobs, 2, msm_pick_window(N, 2), d_ct_out, ctx->stream));
    ctx->launches += msm_last_launches(ctx->ws);
  } else {
    CK(cudaMemsetAsync(d_ct_out, 0, (size_t)2 * G * sizeof(xyzz), ctx->stream));  // ranks >= 2 contribute identities
    if (rank < 2) {
      CK(msm_run(ctx->ws, d_ct_scal, T, d_ct_mont, 2, ct_jobs + rank, 1, msm_pick_window(N, 1), d_ct_out + 2 * rank, ctx->stream));
      ctx->launches += msm_last_launches(ctx->ws);
    }
    int32_t rcg = comm_allgather(ctx, d_ct_out, 2 * sizeof(xyzz), ctx->stream);  // [rank 0: job 0 | rank 1: job 1 | ...]
    if (rcg != MP_OK) return rcg;
  }

  CK(cudaStreamWaitEvent(ctx->stream, S->ev_join, 0));  // join: G1 results are ready for the copies below

  // ---- 6. collect: [bstar | G1 results | CT results | bad flag]
  const size_t res_bytes = sizeof(fr) + (size_t)(J + 4) * sizeof(xyzz) + 16;
  uint8_t* h_res = pinned(S, res_bytes);
  if (!h_res) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  CK(cudaMemcpyAsync(h_res, d_bstar, sizeof(fr), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_res + sizeof(fr), d_g1_out, (size_t)J * sizeof(xyzz), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_res + sizeof(fr) + (size_t)J * sizeof(xyzz), d_ct_out, 4 * sizeof(xyzz), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_res + sizeof(fr) + (size_t)(J + 4) * sizeof(xyzz), d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(stream_wait(ctx, ctx->stream));
  int bad;
  memcpy(&bad, h_res + sizeof(fr) + (size_t)(J + 4) * sizeof(xyzz), sizeof(int));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck or proof point is not a canonical point of the Stark curve");
  fr bstar;
  memcpy(&bstar, h_res, sizeof(fr));
  const xyzz* res = reinterpret_cast<const xyzz*>(h_res + sizeof(fr));
  auto is_id = [&](int j) { return xyzz_is_identity(res[j]); };

  // ---- 7. verdict
  bool g1_id[kG1Checks];
  for (int j = 0; j < kG1Checks; j++) g1_id[j] = is_id(j);
  return verdict(hc, bstar, g1_id, is_id(J) && is_id(J + 1) && is_id(J + 2) && is_id(J + 3));
}

// ------------------------------------------------------------------------------------------
// batched verify (BASELINE config "batch of independent 52-card proofs"): lockstep over B
// proofs -- host threads derive the transcripts and the O(N) scalars of each proof, then ONE
// ciphertext MSM launch sequence (4 jobs per proof) and ONE G1 launch sequence (8 jobs per proof)
// evaluate every group equation of the whole sub-batch.
// ------------------------------------------------------------------------------------------
// flags[g * ncomp + comp] = (sum of the group's job outputs is the identity)
__global__ void __launch_bounds__(64) k_group_identity(const xyzz* __restrict__ outs, int ncomp, int jobs_per_group,
                                                       uint64_t ngroups, uint8_t* __restrict__ flags) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups * ncomp) return;
  uint64_t group = g / ncomp;
  int comp = (int)(g % ncomp);
  xyzz acc = outs[(group * jobs_per_group) * ncomp + comp];
  for (int j = 1; j < jobs_per_group; j++) {
    xyzz v = outs[(group * jobs_per_group + j) * ncomp + comp];
    xyzz_add(acc, v);
  }
  flags[g] = xyzz_is_identity(acc) ? 1 : 0;
}

static int32_t verify_sub_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* decks2,
                                const uint8_t* proofs, size_t Bs, int32_t* statuses, int threads) {
  NvtxRange nvtx("verify_sub_batch");
  ShuffleState* S = ctx->shuffle;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  const Layout L(m, n);
  const size_t plen = shuffle_proof_len(m, n);
  const size_t T1 = 8 * (size_t)m + 5 * (size_t)n + 19;  // G1 terms per proof (see append_g1_checks)
  const size_t SM = 2 * (size_t)m + 3;                   // small ciphertext entries per proof
  const size_t ct_total = 2 * Bs * N + Bs * SM;
  // host staging
  std::vector<uint8_t> g1_pts(Bs * T1 * 64), sm_pts(Bs * SM * 128);
  std::vector<uint32_t> g1_scal(Bs * T1 * 8), ct_scal(ct_total * 8);
  std::vector<HostChecks> hcs(Bs);
  std::vector<fr> bstars(Bs);
  std::vector<int> bad_layout(Bs, 0), malformed(Bs, 0);
  auto work = [&](size_t p) {
    const uint8_t* deck = decks + p * N * 128;
    const uint8_t* deck2 = decks2 + p * N * 128;
    const uint8_t* proof = proofs + p * plen;
    malformed[p] = !proof_scalars_canonical(proof, L);  // reported per item; the group work below still runs
    const Challenges ch = derive_challenges(S, pk, deck, deck2, N, proof, L);
    TermList tl;
    append_g1_checks(tl, S, proof, L, ch, &hcs[p]);
    if (tl.count() != T1) { bad_layout[p] = 1; return; }
    memcpy(&g1_pts[p * T1 * 64], tl.pts.data(), T1 * 64);
    memcpy(&g1_scal[p * T1 * 8], tl.scal.data(), T1 * 32);
    build_ct_plan(S, pk, proof, L, ch, &ct_scal[(p * N) * 8], &ct_scal[(Bs * N + p * N) * 8], &ct_scal[(2 * Bs * N + p * SM) * 8],
                  &sm_pts[p * SM * 128], &bstars[p]);
  };
  S->pool.run(Bs, Bs < 4 ? 1 : threads, work);
  for (size_t p = 0; p < Bs; p++)
    if (bad_layout[p]) return ctx->fail(MP_ERR_INVALID_ARG, "internal: unexpected verifier term count");

  // device buffers
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, ct_total * 128);
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, ct_total * 2 * sizeof(affine));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, ct_total * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, Bs * 8 * sizeof(xyzz));
  uint8_t* d_g1_canon = (uint8_t*)ctx->scratch(sG1Canon, Bs * T1 * 64);
  affine* d_g1_mont = (affine*)ctx->scratch(sG1Mont, Bs * T1 * sizeof(affine));
  uint32_t* d_g1_scal = (uint32_t*)ctx->scratch(sG1Scal, Bs * T1 * 32);
  xyzz* d_g1_out = (xyzz*)ctx->scratch(sG1Out, Bs * kG1Checks * sizeof(xyzz));
  uint8_t* d_flags = (uint8_t*)ctx->scratch(sResults, Bs * 12 + 64);
  int* d_bad = (int*)ctx->scratch(sBadItems, Bs * sizeof(int) + 64);  // one flag per proof: a point of it is off the curve
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_ct_scal); NEED(d_ct_out); NEED(d_g1_canon); NEED(d_g1_mont);
  NEED(d_g1_scal); NEED(d_g1_out); NEED(d_flags); NEED(d_bad);
  cudaStream_t st = ctx->stream;
  CK(cudaMemsetAsync(d_bad, 0, Bs * sizeof(int), st));
  CK(cudaMemcpyAsync(d_ct_canon, decks, Bs * N * 128, cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(d_ct_canon + Bs * N * 128, decks2, Bs * N * 128, cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(d_ct_canon + 2 * Bs * N * 128, sm_pts.data(), sm_pts.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_ct_scal, ct_scal.data(), ct_scal.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_g1_canon, g1_pts.data(), g1_pts.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_g1_scal, g1_scal.data(), g1_scal.size() * 4, cudaMemcpyHostToDevice, st));
  // arena = decks | shuffled decks | per-proof small entries: three runs with their own points-per-proof
  CK(points_to_mont_items((const uint32_t*)d_ct_canon, d_ct_mont, Bs * N * 2, d_bad, N * 2, st));
  CK(points_to_mont_items((const uint32_t*)(d_ct_canon + Bs * N * 128), d_ct_mont + Bs * N * 2, Bs * N * 2, d_bad, N * 2, st));
  CK(points_to_mont_items((const uint32_t*)(d_ct_canon + 2 * Bs * N * 128), d_ct_mont + 4 * Bs * N, Bs * SM * 2, d_bad, SM * 2, st));
  CK(points_to_mont_items((const uint32_t*)d_g1_canon, d_g1_mont, Bs * T1, d_bad, T1, st));
  ctx->launches += 4;
  // ciphertext jobs: per proof (deck, E_m) -> group 0, (deck', small tail) -> group 1
  std::vector<MsmJob> jobs(Bs * 4);
  for (size_t p = 0; p < Bs; p++) {
    const uint32_t a = (uint32_t)(p * N), b = (uint32_t)(2 * Bs * N + p * SM), c2 = (uint32_t)(Bs * N + p * N);
    jobs[4 * p + 0] = MsmJob{a, a, (uint32_t)N};
    jobs[4 * p + 1] = MsmJob{b, b, 1};
    jobs[4 * p + 2] = MsmJob{c2, c2, (uint32_t)N};
    jobs[4 * p + 3] = MsmJob{b + 1, b + 1, (uint32_t)(2 * m + 2)};
  }
  CK(msm_run(ctx->ws, d_ct_scal, ct_total, d_ct_mont, 2, jobs.data(), (int)jobs.size(), msm_pick_window(N / 2 + 1, jobs.size()), d_ct_out, st));
  ctx->launches += msm_last_launches(ctx->ws);
  k_group_identity<<<(unsigned)((Bs * 4 + 63) / 64), 64, 0, st>>>(d_ct_out, 2, 2, Bs * 2, d_flags);
  // G1 jobs: 8 per proof, contiguous terms
  std::vector<MsmJob> g1jobs(Bs * kG1Checks);
  {
    // per-proof job boundaries are identical: take them from a dry layout
    const uint32_t lens[kG1Checks] = {4u, (uint32_t)(2 * m + n + 1), (uint32_t)(m + n + 2), (uint32_t)(2 * m + 3),
                                      (uint32_t)(n + 3), (uint32_t)(n + 2), (uint32_t)(m + n + 2), (uint32_t)(2 * m + 2)};
    for (size_t p = 0; p < Bs; p++) {
      uint32_t off = (uint32_t)(p * T1);
      for (int j = 0; j < kG1Checks; j++) {
        g1jobs[p * kG1Checks + j] = MsmJob{off, off, lens[j]};
        off += lens[j];
      }
    }
  }
  CK(msm_run(ctx->ws, d_g1_scal, Bs * T1, d_g1_mont, 1, g1jobs.data(), (int)g1jobs.size(), msm_pick_window(T1 / kG1Checks, g1jobs.size()), d_g1_out, st));
  ctx->launches += msm_last_launches(ctx->ws);
  k_group_identity<<<(unsigned)((Bs * kG1Checks + 63) / 64), 64, 0, st>>>(d_g1_out, 1, 1, Bs * kG1Checks, d_flags + Bs * 4);
  CK(cudaGetLastError());
  ctx->launches += 2;
  std::vector<uint8_t> flags(Bs * 12);
  std::vector<int> bad(Bs, 0);
  CK(cudaMemcpyAsync(flags.data(), d_flags, Bs * 12, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(bad.data(), d_bad, Bs * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(stream_wait(ctx, st));
  for (size_t p = 0; p < Bs; p++) {
    // a malformed item (point off the curve / not canonical, scalar >= the group order) fails on its own, as the
    // reference's deserialiser would fail it, and does not take the rest of the batch with it
    if (bad[p] || malformed[p]) { statuses[p] = MP_VERIFY_MALFORMED; continue; }
    bool g1_id[kG1Checks];
    for (int j = 0; j < kG1Checks; j++) g1_id[j] = flags[Bs * 4 + p * kG1Checks + j] != 0;
    const uint8_t* cf = &flags[p * 4];
    statuses[p] = verdict(hcs[p], bstars[p], g1_id, cf[0] && cf[1] && cf[2] && cf[3]);
  }
  return MP_OK;
}

int32_t shuffle_verify_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* decks2,
                             const uint8_t* proofs, uint64_t B, int32_t* statuses, int32_t host_threads,
                             const void* d_decks, const void* d_decks2) {
  if (!ctx || !pk || (B && (!decks || !decks2 || !proofs || !statuses))) return MP_ERR_INVALID_ARG;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n, plen = shuffle_proof_len(m, n);
  if (N > small_deck_max()) {
    // large decks: the single-proof verifier per deck, on a few worker contexts so that one
    // proof's serial statement hash (host) overlaps the other proofs' MSMs (device)
    // worker contexts sleep while their kernels run, so their number is not the caller's thread budget; the budget
    // decides whether each worker hashes its own statement or the hashes are shared (StatementHashes)
    const int P = (int)std::max<uint64_t>(1, std::min<uint64_t>(host_threads == 1 ? 1 : 8, B));
    StatementHashes hashes;
    const bool shared = share_statement_hashes(host_threads, P);
    if (shared) hashes.start(S, pk, decks, decks2, N, B);
    return run_on_workers(ctx, P, B, [&](mp_ctx* w, uint64_t p) {
      int32_t st = shuffle_verify(w, pk, decks + p * N * 128, decks2 + p * N * 128, proofs + p * plen,
                                  d_decks ? (const uint8_t*)d_decks + p * N * 128 : nullptr,
                                  d_decks2 ? (const uint8_t*)d_decks2 + p * N * 128 : nullptr, shared ? &hashes : nullptr, p);
      if (st == MP_ERR_NOT_ON_CURVE || st == MP_ERR_NOT_CANONICAL) st = MP_VERIFY_MALFORMED;  // this item only
      if (st >= 0) statuses[p] = st;
      return st < 0 ? st : MP_OK;
    }, /*sleeping_waits=*/P > 1);
  }
  int threads = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  threads = std::max(1, std::min(threads, 64));
  // sub-batches bounded by the job grid (<= 65535 jobs per launch) and ~2^25 ciphertext terms
  size_t sub = std::min<size_t>(4096, std::max<size_t>(1, ((size_t)1 << 24) / N));
  return run_chunks(ctx, B, sub, [&](mp_ctx* w, uint64_t p0, size_t Bs) {
    return verify_sub_batch(w, pk, decks + p0 * N * 128, decks2 + p0 * N * 128, proofs + p0 * plen, Bs, statuses + p0, threads);
  });
}

// ------------------------------------------------------------------------------------------
// batched prove: B independent shuffle_and_remask calls under the same parameters and key.
// Small-deck proofs are latency-bound (a chain of ~5 dependent MSM launches with a 253-doubling
// fold each), so the batch runs P worker contexts concurrently -- one host thread, CUDA stream
// and workspace each -- and the GPU overlaps their kernels.  Proof i is byte-identical to what
// mp_shuffle_and_remask produces for the same inputs.
// ------------------------------------------------------------------------------------------
}  // namespace mp
