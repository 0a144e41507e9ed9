// cd_shared.cuh -- parameters shared by the two coordinate-descent kernels (cd.cu: general; cd_lpc.cu: separable problems)
#pragma once
#include "common.cuh"

namespace qcqp {

enum { MODE_GRAD = 0, MODE_STRICT = 1, MODE_FRESH = 2, MODE_GRAD_GENERAL = 3 };

struct CdK {
    int num_iters;
    double viol_tol, tol;
    int phase1, mode, refresh_every;
};

int lpc_launch(qcqp_pack* p, const CdK& k, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0, double* dmv,
               qcqp_cd_stats* dstats, cudaStream_t stream);

// cd_lpc2.cu: phase 2 of the separable dense-objective path as a resolver / helper CTA with TMA-staged diagonal blocks
int lpc2_launch(qcqp_pack* p, const CdK& k, int R, qcqp_rng_state* drng, double* dX, const double* G, qcqp_cd_stats* dstats,
                cudaStream_t stream);
size_t lpc2_smem_bytes(int n);
bool lpc2_supported(int n);

// cd_blk.cu: one CTA per restart (sparse forms only); blk_wanted: the dispatch rule of qcqp_cd_improve
bool blk_wanted(const qcqp_pack* p);
int blk_launch(qcqp_pack* p, const CdK& k, const double* dX0, int R, qcqp_rng_state* drng, double* dX, double* df0, double* dmv,
               qcqp_cd_stats* dstats, cudaStream_t stream, int force_threads);

}  // namespace qcqp
