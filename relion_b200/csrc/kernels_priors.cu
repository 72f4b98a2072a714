// relion_b200 — per-particle priors for the weight conversion (sm_100a).  The squared-difference kernels live in
// kernels_coarse.cu and kernels_fine.cu.
//
// Replaces cuda_kernel_diff2_coarse / cuda_kernel_diff2_fine
// (/root/reference/src/acc/cuda/cuda_kernels/diff2.cuh:24-189, 193-332) and their ALTCPU twins
// (src/acc/cpu/cpu_kernels/diff2.h:32-430) with kernels batched over a whole pool of particles:
// no per-particle launch, no host sync, image corrections (pixel_correction, corr_img —
// acc_ml_optimiser_impl.h:1251-1268, acc_helper_functions_impl.h:164-196) applied on the fly.
#include "img_src.cuh"

// ---------------------------------------------------------------------------------------------
// priors: pdf_orientation = log(pdf), zero flags (initOrientations, utilities_impl.h:656-668);
// pdf_offset (acc_ml_optimiser_impl.h:2094-2171)
// ---------------------------------------------------------------------------------------------
__global__ void k_prep_priors(const RbPartMeta *metas, RbModelDev M, RbSamplingDev S,
                              const int *dir_idx, const double *dir_prior, const int *psi_idx, const double *psi_prior,
                              float *pdf_orient, unsigned char *pdf_orient_zero,
                              float *pdf_offset, unsigned char *pdf_offset_zero, RbPartState *states)
{
	const int p = blockIdx.y;
	const RbPartMeta m = metas[p];
	const int no = m.nd * m.np;
	const int ndense = M.nr_classes * no;
	for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < ndense; o += gridDim.x * blockDim.x)
	{
		int k = o / no, oi = o - k * no, idl = oi / m.np, ipl = oi - idl * m.np;
		double pdf;
		if (M.do_skip_rotate) pdf = M.pdf_class[k];                                        // :1966-1967
		else if (m.dir_off < 0) pdf = M.pdf_direction[(size_t) k * S.n_dir + idl];
		else pdf = dir_prior[m.dir_off + idl] * psi_prior[m.psi_off + ipl];
		if (!(M.pdf_class[k] > 0.)) pdf = 0.;   // classes with zero pdf_class are never evaluated (:1069)
		pdf_orient_zero[m.prior_off + o] = (pdf == 0.);
		pdf_orient[m.prior_off + o] = (pdf == 0.) ? 0.f : (float) log(pdf);
	}
	if (blockIdx.x == 0)
	{
		// one block per class when the references are 2D and carry their own prior centre (:2100-2104), else one block
		const int Kp = M.prior_classes();
		for (int kt = threadIdx.x; kt < Kp * S.n_trans; kt += blockDim.x)
		{
			const int k = kt / S.n_trans, t = kt - k * S.n_trans;
			const double prx = M.prior_offset_class ? M.prior_offset_class[2 * k] : m.prx;
			const double pry = M.prior_offset_class ? M.prior_offset_class[2 * k + 1] : m.pry;
			double offx = m.oldx + S.trans_x[t], offy = m.oldy + S.trans_y[t];
			double tdiff2 = (offx - prx) * (offx - prx) / (-2. * M.s2off) + (offy - pry) * (offy - pry) / (-2. * M.s2off);
			tdiff2 *= M.pixel_size * M.pixel_size;
			double pdf; bool z;
			if (M.s2off < 0.0001) { z = tdiff2 > 0.; pdf = z ? 0. : 1.; }
			else { z = false; pdf = tdiff2; }
			pdf_offset_zero[(size_t) p * Kp * S.n_trans + kt] = z;
			pdf_offset[(size_t) p * Kp * S.n_trans + kt] = (float) pdf;
		}
		if (threadIdx.x == 0)
		{
			RbPartState st;
			memset(&st, 0, sizeof(st));
			st.min_diff2_bits = 0x7f7fffff; st.fmin_bits = 0x7f7fffff;
			states[p] = st;
		}
	}
}

int rbk_prep_priors(rb_ctx *ctx, PoolSlot &s)
{
	dim3 grid((s.max_no * ctx->d_model.nr_classes + 255) / 256, s.P);
	if (grid.x > 64) grid.x = 64;
	if (grid.x < 1) grid.x = 1;
	k_prep_priors<<<grid, 256, 0, ctx->stream>>>(s.meta.as<RbPartMeta>(), ctx->d_model, ctx->d_samp,
		s.dir_idx.as<int>(), s.dir_prior.as<double>(), s.psi_idx.as<int>(), s.psi_prior.as<double>(),
		s.pdf_orient.as<float>(), s.pdf_orient_zero.as<unsigned char>(),
		s.pdf_offset.as<float>(), s.pdf_offset_zero.as<unsigned char>(), s.state.as<RbPartState>());
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}
