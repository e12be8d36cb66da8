// Device-resident preconditioned conjugate gradient for (tau A^T A + gam2 I) mu = v
// (vamp::precondCG_solver + vamp::lmmse_mult, src/vamp.cpp:645-757).
//
// One CG iteration is a fixed sequence of launches on the context stream:
//     k_ax_partial, k_ax_reduce [, all-reduce N, k_scale_div]      A p         (matrix pass 1)
//     k_atx                                                        A^T (A p)   (matrix pass 2)
//     k_cg_dp      [, all-reduce 1]                                d = tau*.. + gam2 p, <d,p>
//     k_cg_step    [, all-reduce 3]                                mu, r, z, <v,mu>, <r,z>, <r,r>
//     k_cg_finish                                                  stopping tests, beta, p
// alpha, beta and both stopping tests are evaluated by the kernels from device memory; once CgScalars::done is set
// every later launch returns at its first instruction. The host never waits for a dot product: it keeps
// `cg_depth` iterations enqueued ahead and only polls the done flag of an iteration that has already finished.
#include <string.h>
#include <vector>
#include "common.h"

extern "C" int vampomi_cg_solve(vampomi_ctx* c, int rhs_vec, int sol_vec, int warm_start, double tau, double gam2,
                                double tol, int max_iter, int onsager_mode, int* iters, double* rel_err,
                                double* rhs_dot_sol) {
    using namespace vampomi;
    VO_ARG(c && is_mvec(rhs_vec) && is_mvec(sol_vec) && rhs_vec != sol_vec, "cg_solve: rhs and sol must be distinct M-vectors");
    VO_ARG(rhs_vec < VAMPOMI_V_TMP_M0 && sol_vec < VAMPOMI_V_TMP_M0, "cg_solve: work vectors cannot be rhs or sol");
    VO_ARG(max_iter >= 1, "cg_solve: max_iter must be >= 1");
    if (!c->stats_ready) { set_error("cg_solve before compute_stats"); return VAMPOMI_ERR_STATE; }
    VO_CUDA(cudaSetDevice(c->device));

    const double* v = vec_ptr(c, rhs_vec);
    double* mu = vec_ptr(c, sol_vec);
    double* atx_out = c->mvec[VAMPOMI_V_TMP_M0];
    double* tmpN = c->nvec[VAMPOMI_V_TMP_N0 - 32];
    double* p = c->mvec[VAMPOMI_V_CG_P];
    const int* done = &c->cg->done;
    const double diag = tau * (c->N - 1) / c->N + gam2;                      // src/vamp.cpp:676-677

    // r = v - Q mu_start (two matrix passes) unless the start is the zero vector (src/vamp.cpp:647,681-684)
    if (warm_start) {
        VO_CHECK(launch_ax(c, mu, tmpN, nullptr));
        VO_CHECK(launch_atx(c, tmpN, atx_out, nullptr));
    }
    VO_CHECK(launch_cg_init(c, v, mu, atx_out, warm_start ? 1 : 0, tau, gam2, diag, c->sums));
    const bool nccl_scalars = !c->xchg.enabled;          // with the peer-memory exchange the kernels' last block already summed over GPUs
    if (nccl_scalars) VO_CHECK(allreduce_inplace(c, c->sums, 2));
    VO_CHECK(launch_cg_init_finish(c, c->sums));

    int depth = c->tune.cg_depth;
    if (depth < 1) depth = 1;
    if (depth > 32) depth = 32;
    cudaEvent_t ev[32];
    for (int k = 0; k < depth; k++) VO_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    int rc = VAMPOMI_OK;
    // bookkeeping of what was enqueued per CG iteration, so that look-ahead launches that turn out to be no-ops (the
    // done flag was already set when they ran) are not reported as matrix passes / streamed bytes / timed launches
    int launched = 0;
    std::vector<size_t> span_mark;
    if (c->prof_pending.size() > 2048) VO_CHECK(prof_resolve(c));
    const long long pass_bytes = (long long)c->M * c->N * 8;
    for (int i = 0; i < max_iter; i++) {
        const int parity = i & 1, slot = i % depth;
        if (i >= depth) {                                   // poll the flag of iteration i - depth (already retired or close to)
            if (cudaEventSynchronize(ev[slot]) != cudaSuccess) { set_error("cg_solve: event sync failed"); rc = VAMPOMI_ERR_CUDA; break; }
            if (c->cg_poll_host[slot] != 0) break;
        }
        span_mark.push_back(c->prof_pending.size());
        launched = i + 1;
        if ((rc = launch_ax(c, p, tmpN, done)) != VAMPOMI_OK) break;
        if ((rc = launch_atx(c, tmpN, atx_out, done)) != VAMPOMI_OK) break;
        if ((rc = launch_cg_dp(c, atx_out, tau, gam2, c->sums)) != VAMPOMI_OK) break;
        if (nccl_scalars && (rc = allreduce_inplace(c, c->sums, 1)) != VAMPOMI_OK) break;
        if ((rc = launch_cg_step(c, v, mu, diag, parity, c->sums, c->sums + 1)) != VAMPOMI_OK) break;
        if (nccl_scalars && (rc = allreduce_inplace(c, c->sums + 1, 3)) != VAMPOMI_OK) break;
        if ((rc = launch_cg_finish(c, parity, gam2, tol, max_iter, onsager_mode, c->sums)) != VAMPOMI_OK) break;
        if (cudaMemcpyAsync(&c->cg_poll_host[slot], done, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaEventRecord(ev[slot], c->stream) != cudaSuccess) {
            set_error("cg_solve: could not enqueue the completion poll");
            rc = VAMPOMI_ERR_CUDA;
            break;
        }
    }
    CgScalars fin;
    if (rc == VAMPOMI_OK) {
        if (cudaMemcpyAsync(c->sums_host, c->cg, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            set_error("cg_solve: final read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = VAMPOMI_ERR_CUDA;
        } else {
            memcpy(&fin, c->sums_host, sizeof(CgScalars));
            if (fin.iters < launched) {
                const int idle = launched - fin.iters;
                c->counters[1] -= 2LL * idle;
                c->counters[2] -= 2LL * idle * pass_bytes;
                if (c->profile && (size_t)fin.iters < span_mark.size() && span_mark[fin.iters] <= c->prof_pending.size()) {
                    // drop the spans of the idle iterations
                    for (size_t k = span_mark[fin.iters]; k < c->prof_pending.size(); k++) {
                        c->prof_free.push_back(c->prof_pending[k].e0);
                        c->prof_free.push_back(c->prof_pending[k].e1);
                    }
                    c->prof_pending.resize(span_mark[fin.iters]);
                }
            }
            if (iters) *iters = fin.iters;
            if (rel_err) *rel_err = fin.rel_err;
            if (rhs_dot_sol) *rhs_dot_sol = fin.vmu;
        }
    } else {
        cudaStreamSynchronize(c->stream);
    }
    for (int k = 0; k < depth; k++) cudaEventDestroy(ev[k]);
    return rc;
}
