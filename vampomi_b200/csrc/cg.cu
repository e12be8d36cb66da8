// Device-resident preconditioned conjugate gradient for (tau A^T A + gam2 I) mu = v
// (vamp::precondCG_solver + vamp::lmmse_mult, src/vamp.cpp:645-757), for one system or for two systems in lock-step.
//
// One CG iteration is a fixed sequence of launches on the context stream:
//     A p          k_ax_partial + k_ax_reduce              (one system)      matrix pass 1
//                  k_ax_multi + k_ax_reduce_multi          (two systems: ONE read of the marker block for both)
//     A^T (A p)    k_atx / k_atx_smem + k_atx_reduce                          matrix pass 2
//     k_cg_dp      d = tau*.. + gam2 p, <d,p>              [cross-GPU sum inside the kernel, or all-reduce]
//     k_cg_step    mu, r, z, <v,mu>, <r,z>, <r,r>          [same]
//     k_cg_finish  stopping tests, beta, p
// alpha, beta and both stopping tests are evaluated by the kernels from device memory; once a system's CgScalars::done
// is set every later launch skips it (and returns at its first instruction when all systems are done). The host never
// waits for a dot product: it keeps `cg_depth` iterations enqueued ahead and only polls the done flags of an iteration
// that has already finished.
//
// Why two systems: the LMMSE solve (src/vamp.cpp:308-311) and the Onsager solve (:494-501) of a VAMP iteration have the
// same operator and independent right-hand sides. Every pass is bound by streaming A from HBM, so advancing them together
// costs max(k1, k2) instead of k1 + k2 iterations' worth of traffic; each system keeps exactly its own scalars, tests and
// iteration count.
#include <string.h>
#include <vector>
#include "common.h"

namespace vampomi {

struct CgExtra { const double* x = nullptr; double* out = nullptr; };    // out = A x on the first A p pass of a paired solve

static int cg_run(vampomi_ctx* c, CgBatch& b, const int* warm_ata_given, double tau, double gam2, double tol, int max_iter,
                  CgExtra extra, int* iters, double* rel_err, double* rhs_dot_sol) {
    const int S = b.S;
    const double diag = tau * (c->N - 1) / c->N + gam2;                      // src/vamp.cpp:676-677
    const bool nccl_scalars = !c->xchg.enabled;          // with the peer-memory exchange the kernels' last block already summed over GPUs
    double* sums_dp = c->sums;                           // [S]
    double* sums_st = c->sums + 4;                       // [3S] (init: [2S])

    // r = v - Q mu_start unless the start is the zero vector (src/vamp.cpp:647,681-684): two matrix passes per warm system
    // whose A^T A mu_start the caller does not already hold
    for (int s = 0; s < S; s++) {
        if (b.s[s].warm && !warm_ata_given[s]) {
            VO_CHECK(launch_ax(c, b.s[s].mu, b.s[s].tmpN, nullptr));
            VO_CHECK(launch_atx(c, b.s[s].tmpN, b.s[s].atx_out, nullptr));
        }
    }
    VO_CHECK(launch_cg_init(c, b, tau, gam2, diag, sums_st));
    if (nccl_scalars) VO_CHECK(allreduce_inplace(c, sums_st, 2 * S));
    VO_CHECK(launch_cg_init_finish(c, b, sums_st));
    for (int s = 0; s < S; s++) b.s[s].atx_out = b.s[s].atx_work;            // from here on: A^T A p of the current iteration

    // One-pass CG (knob cg_onepass, schedule "onepass"; kernels_gram.cu): q = A p is a vector of its own. It starts as A p_0
    // (the only A x pass of the solve — it also carries the caller's extra product), A r starts as diag * q (p_0 = r_0/diag),
    // and from then on ONE fused pass per iteration delivers t = A^T q (= A^T A p) and w = A t, from which k_cg_step /
    // k_cg_finish advance A r and q next to r and p. Every `refresh` iterations q and A r are recomputed from p and z by a
    // pass of their own, which bounds the drift of the recurrences in long solves.
    const bool onepass = c->tune.cg_onepass != 0 && gram_supported(c);
    const int refresh = c->tune.gram_refresh > 0 ? c->tune.gram_refresh : 32;
    if (onepass) {
        MultiVec mv{};
        mv.K = S;
        for (int s = 0; s < S; s++) {
            b.s[s].gw = c->nvec[(s == 0 ? VAMPOMI_V_GRAM_W0 : VAMPOMI_V_GRAM_W1) - 32];
            b.s[s].gar = c->nvec[(s == 0 ? VAMPOMI_V_GRAM_AR0 : VAMPOMI_V_GRAM_AR1) - 32];
            mv.in[s] = b.s[s].p; mv.out[s] = b.s[s].tmpN; mv.done[s] = nullptr;
        }
        if (extra.x) { mv.in[S] = extra.x; mv.out[S] = extra.out; mv.done[S] = nullptr; mv.K = S + 1; }
        VO_CHECK(launch_ax_multi(c, mv));
        for (int s = 0; s < S; s++) VO_CHECK(launch_lincomb(c, b.s[s].gar, diag, b.s[s].tmpN, 0.0, b.s[s].tmpN, 1.0, c->N));
    }
    int depth = c->tune.cg_depth;
    if (depth < 1) depth = 1;
    if (depth > 32) depth = 32;
    if (nccl_scalars && c->nranks > 1) depth = 1;        // library collectives on stale buffers are not free: no look-ahead launches there
    cudaEvent_t* ev = c->cg_events;                      // created once per context, reused by every solve
    for (int k = 0; k < depth; k++)
        if (!ev[k]) VO_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    int rc = VAMPOMI_OK;
    // bookkeeping of what was enqueued per CG iteration, so that look-ahead launches that turn out to be no-ops (the
    // done flags were already set when they ran) are not reported as matrix passes / streamed bytes / timed launches
    int launched = 0;
    std::vector<size_t> span_mark;
    std::vector<long long> pass_mark;                       // matrix-pass counter at the start of every enqueued iteration
    if (c->prof_pending.size() > 2048) VO_CHECK(prof_resolve(c));
    const long long pass_bytes = (long long)c->M * c->N * c->elem_bytes;
    const int* done0 = &b.s[0].cg->done;
    for (int i = 0; i < max_iter; i++) {
        const int parity = i & 1, slot = i % depth;
        if (i >= depth) {                                   // poll the flags of iteration i - depth (already retired or close to)
            if (cudaEventSynchronize(ev[slot]) != cudaSuccess) { set_error("cg_solve: event sync failed"); rc = VAMPOMI_ERR_CUDA; break; }
            bool all_done = true;
            for (int s = 0; s < S; s++) all_done = all_done && c->cg_poll_host[2 * slot + s] != 0;
            if (all_done) break;
        }
        span_mark.push_back(c->prof_pending.size());
        pass_mark.push_back(c->counters[1]);
        launched = i + 1;
        if (onepass) {
            if (i > 0 && i % refresh == 0) {                // q = A p and A r = diag * A z afresh (one pass for all of them)
                MultiVec mv{};
                mv.K = 2 * S;
                for (int s = 0; s < S; s++) {
                    mv.in[2 * s] = b.s[s].p; mv.out[2 * s] = b.s[s].tmpN; mv.done[2 * s] = &b.s[s].cg->done;
                    mv.in[2 * s + 1] = b.s[s].z; mv.out[2 * s + 1] = b.s[s].gar; mv.done[2 * s + 1] = &b.s[s].cg->done;
                }
                if ((rc = launch_ax_multi(c, mv)) != VAMPOMI_OK) break;
                for (int s = 0; s < S && rc == VAMPOMI_OK; s++) rc = launch_scale_div(c, b.s[s].gar, b.s[s].gar, 1.0 / diag, c->N, &b.s[s].cg->done);
                if (rc != VAMPOMI_OK) break;
            }
            MultiVec mq{};
            mq.K = S;
            double* w_out[2] = {nullptr, nullptr};
            for (int s = 0; s < S; s++) { mq.in[s] = b.s[s].tmpN; mq.out[s] = b.s[s].atx_out; mq.done[s] = &b.s[s].cg->done; w_out[s] = b.s[s].gw; }
            if ((rc = launch_gram(c, mq, w_out)) != VAMPOMI_OK) break;
        } else if (S == 1 && !(i == 0 && extra.x)) {
            if ((rc = launch_ax(c, b.s[0].p, b.s[0].tmpN, done0)) != VAMPOMI_OK) break;
            if ((rc = launch_atx(c, b.s[0].tmpN, b.s[0].atx_out, done0)) != VAMPOMI_OK) break;
        } else {
            MultiVec mv{};
            mv.K = S;
            for (int s = 0; s < S; s++) { mv.in[s] = b.s[s].p; mv.out[s] = b.s[s].tmpN; mv.done[s] = &b.s[s].cg->done; }
            if (i == 0 && extra.x) { mv.in[S] = extra.x; mv.out[S] = extra.out; mv.done[S] = nullptr; mv.K = S + 1; }
            if ((rc = launch_ax_multi(c, mv)) != VAMPOMI_OK) break;
            MultiVec mt{};
            mt.K = S;
            for (int s = 0; s < S; s++) { mt.in[s] = b.s[s].tmpN; mt.out[s] = b.s[s].atx_out; mt.done[s] = &b.s[s].cg->done; }
            if ((rc = launch_atx_multi(c, mt)) != VAMPOMI_OK) break;
        }
        if ((rc = launch_cg_dp(c, b, tau, gam2, sums_dp)) != VAMPOMI_OK) break;
        if (nccl_scalars && (rc = allreduce_inplace(c, sums_dp, S)) != VAMPOMI_OK) break;
        if ((rc = launch_cg_step(c, b, tau, gam2, diag, parity, sums_dp, sums_st)) != VAMPOMI_OK) break;
        if (nccl_scalars && (rc = allreduce_inplace(c, sums_st, 3 * S)) != VAMPOMI_OK) break;
        if ((rc = launch_cg_finish(c, b, parity, gam2, diag, tol, max_iter, sums_st)) != VAMPOMI_OK) break;
        bool ok = true;
        for (int s = 0; s < S; s++)
            ok = ok && cudaMemcpyAsync(&c->cg_poll_host[2 * slot + s], &b.s[s].cg->done, sizeof(int), cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
        if (!ok || cudaEventRecord(ev[slot], c->stream) != cudaSuccess) {
            set_error("cg_solve: could not enqueue the completion poll");
            rc = VAMPOMI_ERR_CUDA;
            break;
        }
    }
    if (rc == VAMPOMI_OK) {
        // c->cg holds the (up to) two CgScalars back to back
        if (cudaMemcpyAsync(c->sums_host, c->cg, 2 * sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            set_error("cg_solve: final read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = VAMPOMI_ERR_CUDA;
        } else {
            CgScalars fin[2];
            memcpy(fin, c->sums_host, 2 * sizeof(CgScalars));
            int ran = 0;                                    // iterations in which at least one system was still active
            for (int s = 0; s < S; s++) ran = fin[s].iters > ran ? fin[s].iters : ran;
            if (ran < launched) {
                const long long idle_passes = c->counters[1] - pass_mark[ran];     // passes enqueued by the idle iterations
                c->counters[1] -= idle_passes;
                c->counters[2] -= idle_passes * pass_bytes;
                if (c->profile && (size_t)ran < span_mark.size() && span_mark[ran] <= c->prof_pending.size()) {
                    // drop the spans of the idle iterations
                    for (size_t k = span_mark[ran]; k < c->prof_pending.size(); k++) {
                        c->prof_free.push_back(c->prof_pending[k].e0);
                        c->prof_free.push_back(c->prof_pending[k].e1);
                    }
                    c->prof_pending.resize(span_mark[ran]);
                }
            }
            for (int s = 0; s < S; s++) {
                if (iters) iters[s] = fin[s].iters;
                if (rel_err) rel_err[s] = fin[s].rel_err;
                if (rhs_dot_sol) rhs_dot_sol[s] = fin[s].vmu;
            }
        }
    } else {
        cudaStreamSynchronize(c->stream);
    }
    if (rc == VAMPOMI_OK) rc = xchg_check(c);
    return rc;
}

static void fill_sys(vampomi_ctx* c, CgSys* q, int s, int rhs_vec, int sol_vec, int warm, int onsager_mode) {
    q->v = vec_ptr(c, rhs_vec);
    q->mu = vec_ptr(c, sol_vec);
    q->r = c->mvec[s == 0 ? VAMPOMI_V_CG_R : VAMPOMI_V_CG2_R];
    q->z = c->mvec[s == 0 ? VAMPOMI_V_CG_Z : VAMPOMI_V_CG2_Z];
    q->p = c->mvec[s == 0 ? VAMPOMI_V_CG_P : VAMPOMI_V_CG2_P];
    q->d = c->mvec[s == 0 ? VAMPOMI_V_CG_D : VAMPOMI_V_CG2_D];
    q->atx_work = c->mvec[s == 0 ? VAMPOMI_V_TMP_M0 : VAMPOMI_V_TMP_M1];
    q->atx_out = q->atx_work;
    q->tmpN = c->nvec[(s == 0 ? VAMPOMI_V_TMP_N0 : VAMPOMI_V_TMP_N1) - 32];
    q->cg = c->cg + s;
    q->amu = nullptr;
    q->gw = nullptr; q->gar = nullptr;
    q->warm = warm ? 1 : 0;
    q->onsager_mode = onsager_mode ? 1 : 0;
}

}  // namespace vampomi

extern "C" int vampomi_cg_solve(vampomi_ctx* c, int rhs_vec, int sol_vec, int warm_start, double tau, double gam2,
                                double tol, int max_iter, int onsager_mode, int* iters, double* rel_err,
                                double* rhs_dot_sol) {
    using namespace vampomi;
    VO_ARG(c && is_mvec(rhs_vec) && is_mvec(sol_vec) && rhs_vec != sol_vec, "cg_solve: rhs and sol must be distinct M-vectors");
    VO_ARG(!is_work_mvec(rhs_vec) && !is_work_mvec(sol_vec), "cg_solve: work vectors cannot be rhs or sol");
    VO_ARG(max_iter >= 1, "cg_solve: max_iter must be >= 1");
    if (!c->stats_ready) { set_error("cg_solve before compute_stats"); return VAMPOMI_ERR_STATE; }
    VO_CUDA(cudaSetDevice(c->device));
    CgBatch b{};
    b.S = 1;
    fill_sys(c, &b.s[0], 0, rhs_vec, sol_vec, warm_start, onsager_mode);
    const int given[2] = {0, 0};
    return cg_run(c, b, given, tau, gam2, tol, max_iter, CgExtra{}, iters, rel_err, rhs_dot_sol);
}

extern "C" int vampomi_cg_solve_pair(vampomi_ctx* c, const int rhs_vec[2], const int sol_vec[2], const int warm_start[2],
                                     const int warm_ata_vec[2], double tau, double gam2, double tol, int max_iter,
                                     const int onsager_mode[2], int extra_x_vec, int extra_out_vec, const int track_ax_vec[2],
                                     int iters[2], double rel_err[2], double rhs_dot_sol[2]) {
    using namespace vampomi;
    VO_ARG(c && rhs_vec && sol_vec && warm_start && onsager_mode, "cg_solve_pair: NULL argument");
    for (int s = 0; s < 2; s++) {
        VO_ARG(is_mvec(rhs_vec[s]) && is_mvec(sol_vec[s]) && !is_work_mvec(rhs_vec[s]) && !is_work_mvec(sol_vec[s]),
               "cg_solve_pair: rhs and sol must be non-work M-vectors");
        if (warm_start[s] && warm_ata_vec && warm_ata_vec[s] >= 0)
            VO_ARG(is_mvec(warm_ata_vec[s]) && !is_work_mvec(warm_ata_vec[s]) && warm_ata_vec[s] != sol_vec[s] && warm_ata_vec[s] != rhs_vec[s],
                   "cg_solve_pair: warm_ata_vec must be a non-work M-vector other than rhs/sol");
    }
    VO_ARG(rhs_vec[0] != sol_vec[0] && rhs_vec[1] != sol_vec[1] && sol_vec[0] != sol_vec[1] && rhs_vec[0] != sol_vec[1] && rhs_vec[1] != sol_vec[0],
           "cg_solve_pair: the two solutions and the right-hand sides must be distinct vectors");
    VO_ARG(max_iter >= 1, "cg_solve_pair: max_iter must be >= 1");
    CgExtra extra;
    if (extra_x_vec >= 0) {
        VO_ARG(is_mvec(extra_x_vec) && !is_work_mvec(extra_x_vec) && vec_ptr(c, extra_out_vec) && !is_mvec(extra_out_vec) &&
               extra_out_vec != VAMPOMI_V_TMP_N0 && extra_out_vec != VAMPOMI_V_TMP_N1 && extra_x_vec != sol_vec[0] && extra_x_vec != sol_vec[1],
               "cg_solve_pair: extra needs a non-work M-vector in and a non-work N-vector out");
        extra.x = vec_ptr(c, extra_x_vec);
        extra.out = vec_ptr(c, extra_out_vec);
    }
    if (!c->stats_ready) { set_error("cg_solve_pair before compute_stats"); return VAMPOMI_ERR_STATE; }
    VO_CUDA(cudaSetDevice(c->device));
    CgBatch b{};
    b.S = 2;
    int given[2] = {0, 0};
    for (int s = 0; s < 2; s++) {
        fill_sys(c, &b.s[s], s, rhs_vec[s], sol_vec[s], warm_start[s], onsager_mode[s]);
        if (warm_start[s] && warm_ata_vec && warm_ata_vec[s] >= 0) { b.s[s].atx_out = vec_ptr(c, warm_ata_vec[s]); given[s] = 1; }
        if (track_ax_vec && track_ax_vec[s] >= 0) {
            const int t = track_ax_vec[s];
            VO_ARG(vec_ptr(c, t) && !is_mvec(t) && t != VAMPOMI_V_TMP_N0 && t != VAMPOMI_V_TMP_N1 && t != extra_out_vec &&
                   (s == 0 || !track_ax_vec || t != track_ax_vec[0]),
                   "cg_solve_pair: track_ax_vec must name distinct non-work N-vectors");
            b.s[s].amu = vec_ptr(c, t);
        }
    }
    return cg_run(c, b, given, tau, gam2, tol, max_iter, extra, iters, rel_err, rhs_dot_sol);
}
