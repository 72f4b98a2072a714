/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle of the per-particle E-step driver.
 *
 * Restates accDoExpectationOneParticle (/root/reference/src/acc/acc_ml_optimiser_impl.h:3672-3962)
 * for the supported subset (3D reference, 2D images, one image per particle, nr_bodies == 1,
 * adaptive_oversampling > 0, Gaussian or first-iteration cross-correlation criterion, no helices) on top of a
 * kernel table (oracle_kernels.h), so the
 * same orchestration runs on the compiled reference kernels ("reference") or on the restated ones
 * ("port").  Inputs/outputs use the product's public structs (include/relion_b200.h) so that a test
 * can hand the identical descriptors to rb_estep_pool() and to oracle_estep_pool().
 */
#ifndef ORACLE_ESTEP_DRIVER_H_
#define ORACLE_ESTEP_DRIVER_H_

#include "oracle_kernels.h"
#include "relion_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Optional per-particle dump of intermediate arrays for stage-level parity tests.
 * Buffers are caller-allocated with the stated capacities; NULL pointers are skipped. */
typedef struct {
	int particle;                 /* which particle of the pool to dump                          */
	float *coarse_diff2;          /* [K*nd*np*T] Mweight after the coarse diff2 pass (lowest() = not computed) */
	float *coarse_weights;        /* [K*nd*np*T] after exponentiation                            */
	unsigned char *coarse_significant; /* [K*nd*np*T]                                            */
	int64_t fine_capacity;        /* capacity of the fine_* arrays                               */
	int64_t fine_count;           /* out: number of fine samples                                 */
	int64_t *fine_ihidden_over;   /* [fine_count]                                                */
	float *fine_diff2;            /* [fine_count] before conversion                              */
	float *fine_weights;          /* [fine_count] after exponentiation                           */
	float *wdiff2s_parts;         /* [Np_cur] per-pixel wavg output                              */
	float *wdiff2s_AA, *wdiff2s_XA; /* [K*Np_cur]                                                */
} ok_debug;

/* Runs the pool, one particle per OpenMP task (mirrors ALTCPU's tbb::parallel_for over particles,
 * src/ml_optimiser.cpp:4294-4309).  flags as rb_estep_pool.  exact_threshold: 0 = the reference's
 * sequential fp32 sort+scan threshold (ALTCPU semantics, src/acc/utilities.h:383-398);
 * 1 = same rule evaluated in exact (double) arithmetic — what the sort-free GPU selector computes.
 * Returns 0 or a negative rb_status. */
int oracle_estep_pool(const ok_kernel_table *K, const rb_model *m, const rb_sampling *s,
                      const ok_projector *refs, ok_backprojector *bps,
                      const rb_particles *pool, rb_pool_out *out,
                      unsigned flags, int num_threads, int exact_threshold, ok_debug *dbg);

/* The significance rule alone (findThresholdIdxInCumulativeSum + sort + scan,
 * acc_helper_functions.h:191-240, utilities.h:301-398) on an array of weights.
 * filter_zero: coarse pass (only weights > 0 are sorted).  Returns thresholdIdx; fills outputs. */
int64_t oracle_significance(const float *weights, int64_t n, double adaptive_fraction,
                            int maximum_significants, int filter_zero, int exact,
                            float *sum_weight, float *significant_weight, int64_t *n_filtered);

const ok_kernel_table *portk_kernel_table(void);

#ifdef __cplusplus
}
#endif
#endif
