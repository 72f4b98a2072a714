// relion_b200 — C-ABI glue (include/relion_b200.h): context, model/sampling upload, pool pipeline.
// Product code: no reference to oracle/; fails loudly without a CUDA device (no CPU fallback).
#include "common.cuh"
#include <cstdarg>
#include <nvtx3/nvToolsExt.h>
#include <cstdlib>
#include <cmath>
#include <algorithm>

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void rb_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *rb_last_error(void) { return g_err; }
extern "C" int rb_version(void) { return RB_VERSION; }

int DevBuf::ensure(size_t n)
{
	if (n <= bytes && p) return RB_OK;
	if (n == 0) n = 16;
	if (p) { RB_CUDA(cudaFree(p)); p = nullptr; bytes = 0; }
	RB_CUDA(cudaMalloc(&p, n));
	bytes = n;
	return RB_OK;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }

static int upload(rb_ctx *ctx, DevBuf &b, const void *src, size_t bytes)
{
	RB_CHECK(b.ensure(bytes));
	if (bytes) RB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int rb_ctx_create(int device, rb_ctx **out)
{
	RB_ARG(out != nullptr, "rb_ctx_create: out is NULL");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
	{
		rb_set_error("rb_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
		             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return RB_ERR_CUDA;
	}
	RB_ARG(device >= 0 && device < ndev, "rb_ctx_create: device %d out of range (have %d)", device, ndev);
	RB_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	RB_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
	{
		rb_set_error("rb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
		return RB_ERR_CUDA;
	}
	rb_ctx *ctx = new rb_ctx();
	for (int i = 0; i < RB_MAX_CLASSES; i++) { ctx->gemmA_stamp[i] = -1; ctx->core_stamp[i] = -1; ctx->core_R[i] = 0; }
	ctx->device = device;
	ctx->num_sms = prop.multiProcessorCount;
	RB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	RB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < RB_NUM_SLOTS; i++)
	{
		RB_CUDA(cudaEventCreateWithFlags(&ctx->slot[i].uploaded, cudaEventDisableTiming));
		RB_CUDA(cudaEventCreateWithFlags(&ctx->slot[i].done, cudaEventDisableTiming));
	}
	RB_CUDA(cudaStreamCreateWithFlags(&ctx->fetch_stream, cudaStreamNonBlocking));
	memset(ctx->proj, 0, sizeof(ctx->proj));
	memset(ctx->bp, 0, sizeof(ctx->bp));
	*out = ctx;
	return RB_OK;
}

static void release_slot(PoolSlot &s)
{
	DevBuf *bufs[] = {&s.Fimg, &s.Fnomask, &s.Fctf, &s.meta, &s.state, &s.dir_idx, &s.dir_prior, &s.psi_idx, &s.psi_prior,
	                  &s.Mweight, &s.pdf_orient, &s.pdf_orient_zero, &s.pdf_offset, &s.pdf_offset_zero,
	                  &s.so_list, &s.pair_list, &s.fo, &s.fs_w, &s.fs_ihid, &s.counters, &s.shells, &s.out_pdf_dir, &s.out_pdf_class,
	                  &s.fimg4, &s.cimg4, &s.slices, &s.cc_corr, &s.simg4, &s.sst, &s.sctf, &s.bp_cnt, &s.bp_item_of, &s.bp_items, &s.bp_samp};
	for (DevBuf *b : bufs) b->release();
	if (s.uploaded) cudaEventDestroy(s.uploaded);
	if (s.done) cudaEventDestroy(s.done);
}

extern "C" void rb_ctx_destroy(rb_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	for (int i = 0; i < RB_MAX_CLASSES; i++) { ctx->proj_buf[i].release(); ctx->proj8_buf[i].release(); ctx->proj2_buf[i].release(); ctx->proj2c_buf[i].release(); ctx->proj4c_buf[i].release(); ctx->bp_buf[i].release(); ctx->bp_blk_buf[i].release(); }
	DevBuf *bufs[] = {&ctx->s_coarse_eulers, &ctx->s_over_rot, &ctx->s_over_tilt, &ctx->s_over_psi, &ctx->s_rot, &ctx->s_tilt,
	                  &ctx->s_psi, &ctx->s_ctx, &ctx->s_cty, &ctx->s_ftx, &ctx->s_fty, &ctx->s_tx, &ctx->s_ty, &ctx->s_otx, &ctx->s_oty,
	                  &ctx->m_rows_c, &ctx->m_rows_f, &ctx->m_ires_c, &ctx->m_ires_f,
	                  &ctx->m_cc[0], &ctx->m_cc[1], &ctx->m_cc[2], &ctx->m_cc[3], &ctx->m_cc[4], &ctx->m_cc[5],
	                  &ctx->m_pix_c, &ctx->m_pix_f, &ctx->m_minvs2, &ctx->m_pdf_dir, &ctx->m_pdf_class, &ctx->m_dvp, &ctx->d_proj, &ctx->d_bp,
	                  &ctx->m_pix_rs, &ctx->band_slices, &ctx->band_tabc, &ctx->band_tabo, &ctx->band_tabu, &ctx->comm_buf, &ctx->comm_buf2, &ctx->posed_pix, &ctx->posed_sorted};
	for (DevBuf *b : bufs) b->release();
	for (rb_ctx::BlockTable *t : ctx->blk_tables) { t->buf.release(); delete t; }
	ctx->blk_tables.clear();
	for (auto &b : ctx->scratch) b.release();
	for (auto &b : ctx->grid_buf) b.release();
	for (auto &b : ctx->gemm_buf) b.release();
	for (int i = 0; i < RB_MAX_CLASSES; i++) for (auto &b : ctx->gemmA[i]) b.release();
	for (auto &b : ctx->gemmA_all) b.release();
	for (auto &b : ctx->wc_buf) b.release();
	for (auto &b : ctx->prep_buf) b.release();
	rbk_prepare_release(ctx);
	for (auto &b : ctx->recon_buf) b.release();
	for (int i = 0; i < RB_NUM_SLOTS; i++) for (auto &b : ctx->prep_raw[i]) b.release();
	for (int i = 0; i < 2; i++) { for (auto &b : ctx->posed_buf[i]) b.release(); if (ctx->posed_ev[i]) cudaEventDestroy(ctx->posed_ev[i]); }
	for (int i = 0; i < RB_NUM_SLOTS; i++) release_slot(ctx->slot[i]);
	for (auto &kv : ctx->stage_ev) { cudaEventDestroy(kv.second.first); cudaEventDestroy(kv.second.second); }
	cudaStreamDestroy(ctx->stream);
	cudaStreamDestroy(ctx->copy_stream);
	if (ctx->fetch_stream) cudaStreamDestroy(ctx->fetch_stream);
	if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->aux_fork); cudaEventDestroy(ctx->aux_join); }
	delete ctx;
}

extern "C" int rb_sync(rb_ctx *ctx)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx, "rb_sync: ctx is NULL");
	RB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" long long rb_launch_count(rb_ctx *ctx) { return ctx ? ctx->launches : -1; }

// Stage brackets: a cudaEvent pair per stage (rb_stage_ms) and an NVTX range of the same name on the calling thread, so that a
// timeline tool (nsys / ncu --nvtx) shows the stages the way the reference's CTIC/CTOC timers and its CUDA_PROFILING NVTX
// ranges do (/root/reference/src/acc/cuda/cuda_settings.h, src/acc/acc_ml_optimiser_impl.h LAUNCH_PRIVATE_ERROR / CTIC blocks).
// nvtx3 is header-only and a no-op (one predictable branch) when no tool is attached.
int rb_stage_begin(rb_ctx *ctx, const char *name)
{
	if (name[0] != '_') nvtxRangePushA(name);
	auto it = ctx->stage_ev.find(name);
	if (it == ctx->stage_ev.end())
	{
		cudaEvent_t a, b;
		RB_CUDA(cudaEventCreate(&a)); RB_CUDA(cudaEventCreate(&b));
		it = ctx->stage_ev.emplace(name, std::make_pair(a, b)).first;
	}
	RB_CUDA(cudaEventRecord(it->second.first, ctx->stream));
	return RB_OK;
}
int rb_stage_end(rb_ctx *ctx, const char *name)
{
	auto it = ctx->stage_ev.find(name);
	if (it == ctx->stage_ev.end()) return RB_OK;
	if (name[0] != '_') nvtxRangePop();
	RB_CUDA(cudaEventRecord(it->second.second, ctx->stream));
	return RB_OK;
}
extern "C" double rb_stage_ms(rb_ctx *ctx, const char *stage)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	if (!ctx || !stage) return -1.;
	auto it = ctx->stage_ev.find(stage);
	if (it == ctx->stage_ev.end()) return -1.;
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, it->second.first, it->second.second) != cudaSuccess) { cudaGetLastError(); return -1.; }
	return (double) ms;
}

extern "C" int rb_timer_start(rb_ctx *ctx)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx, "rb_timer_start: ctx is NULL");
	return rb_stage_begin(ctx, "__timer");
}

extern "C" int rb_timer_stop(rb_ctx *ctx, double *ms)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && ms, "rb_timer_stop: NULL argument");
	RB_CHECK(rb_stage_end(ctx, "__timer"));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	*ms = rb_stage_ms(ctx, "__timer");
	return RB_OK;
}

int rb_sync_tables(rb_ctx *ctx)
{
	RB_CHECK(upload(ctx, ctx->d_proj, ctx->proj, sizeof(ctx->proj)));
	RB_CHECK(upload(ctx, ctx->d_bp, ctx->bp, sizeof(ctx->bp)));
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// reference volumes / accumulators
// ---------------------------------------------------------------------------------------------
// Block table of an expanded reference (RbProjector::blk): rank of every 4 x 4 x 4 block when the blocks are ordered by the
// squared distance of their centre from the origin (counting sort on the integer key, ties in index order: deterministic).
// RB_BLOCK_SORT=0 keeps the blocks in (z, y, x) order (A/B).
static int block_table(rb_ctx *ctx, int mdlX, int mdlY, int mdlZ, int initY, int initZ, const uint32_t **table)
{
	const int geom[5] = {mdlX, mdlY, mdlZ, initY, initZ};
	for (rb_ctx::BlockTable *t : ctx->blk_tables)
		if (!memcmp(t->geom, geom, sizeof(geom))) { *table = t->buf.as<uint32_t>(); return RB_OK; }
	const int nbx = (mdlX + 3) / 4, nby = (mdlY + 3) / 4, nbz = (mdlZ + 3) / 4;
	const size_t nb = (size_t) nbx * nby * nbz;
	const int kbx = (nbx + 3) / 4, kby = (nby + 3) / 4, kbz = (nbz + 1) / 2;            // bricks of 4 x 4 x 2 blocks (rb_blk_slot)
	const size_t nslot = (size_t) kbx * kby * kbz * 32;
	std::vector<uint32_t> key(nb), rank(nb), slots(nslot, 0u);
	uint32_t kmax = 0;
	for (int bz = 0; bz < nbz; bz++)
		for (int by = 0; by < nby; by++)
			for (int bx = 0; bx < nbx; bx++)
			{
				const long long cx = 4 * bx + 2, cy = 4 * by + 2 + initY, cz = 4 * bz + 2 + initZ;
				const uint32_t kk = (uint32_t) (cx * cx + cy * cy + cz * cz);
				key[((size_t) bz * nby + by) * nbx + bx] = kk;
				kmax = std::max(kmax, kk);
			}
	const char *e = getenv("RB_BLOCK_SORT");
	if (e && atoi(e) == 0) { for (size_t i = 0; i < nb; i++) rank[i] = (uint32_t) i; }
	else
	{
		std::vector<uint32_t> start((size_t) kmax + 2, 0);
		for (size_t i = 0; i < nb; i++) start[key[i] + 1]++;
		for (size_t i = 1; i < start.size(); i++) start[i] += start[i - 1];
		for (size_t i = 0; i < nb; i++) rank[i] = start[key[i]]++;
	}
	for (int bz = 0; bz < nbz; bz++)
		for (int by = 0; by < nby; by++)
			for (int bx = 0; bx < nbx; bx++)
				slots[rb_blk_slot(kbx, kbx * kby, bx, by, bz)] = rank[((size_t) bz * nby + by) * nbx + bx];
	rb_ctx::BlockTable *t = new rb_ctx::BlockTable();
	memcpy(t->geom, geom, sizeof(geom));
	int rc = t->buf.ensure(nslot * sizeof(uint32_t));
	if (rc != RB_OK) { delete t; return rc; }
	RB_CUDA(cudaMemcpyAsync(t->buf.p, slots.data(), nslot * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->blk_tables.push_back(t);
	*table = t->buf.as<uint32_t>();
	return RB_OK;
}

static int set_reference_common(rb_ctx *ctx, int k, int mdlX, int mdlY, int &mdlZ, int initY, int &initZ, int maxR, double pf)
{
	RB_ARG(ctx, "ctx is NULL");
	RB_ARG(k >= 0 && k < RB_MAX_CLASSES, "class index %d out of range", k);
	RB_ARG(mdlX > 1 && mdlY > 1 && mdlZ >= 1, "rb_set_reference: bad dimensions %dx%dx%d", mdlX, mdlY, mdlZ);
	RB_CUDA(cudaSetDevice(ctx->device));
	ctx->ref_2d[k] = (mdlZ == 1);
	ctx->ref_version[k]++;
	if (mdlZ == 1) { mdlZ = 2; initZ = 0; }             // 2D reference (AccProjector with mdlZ == 0 in the reference): plane 1 stays zero
	size_t n = (size_t) mdlX * mdlY * mdlZ;
	const int nbx = (mdlX + 3) / 4, nby = (mdlY + 3) / 4, nbz = (mdlZ + 3) / 4;
	const size_t ncell = (size_t) nbx * nby * nbz * 64;
	RB_ARG(ncell < ((size_t) 1 << 31), "rb_set_reference: %dx%dx%d voxels exceed the cell index range", mdlX, mdlY, mdlZ);
	const uint32_t *blk = nullptr;
	RB_CHECK(block_table(ctx, mdlX, mdlY, mdlZ, initY, initZ, &blk));
	RB_CHECK(ctx->proj_buf[k].ensure(n * sizeof(float2)));
	RB_CHECK(ctx->proj8_buf[k].ensure(ncell * 4 * sizeof(float4)));
	RB_CHECK(ctx->proj2_buf[k].ensure(n * sizeof(float4)));
	RbProjector &p = ctx->proj[k];
	p.mdl = ctx->proj_buf[k].as<float2>();
	p.mdl8 = ctx->proj8_buf[k].as<float4>();
	p.mdl2 = ctx->proj2_buf[k].as<float4>();
	p.mdlX = mdlX; p.mdlY = mdlY; p.mdlZ = mdlZ; p.mdlXY = mdlX * mdlY;
	p.mdlInitY = initY; p.mdlInitZ = initZ; p.mdlMaxR = maxR; p.padding_factor = (float) pf;
	p.c2X = mdlX; p.c2XY = p.mdlXY; p.c2InitY = initY; p.c2InitZ = initZ;       // full x-pair copy until a pool asks for the core
	p.quad = nullptr;
	p.blk = blk; p.nbx = (nbx + 3) / 4; p.nbxy = p.nbx * ((nby + 3) / 4);       // brick grid of the table (rb_blk_slot)
	ctx->core_stamp[k] = -1;
	ctx->has_proj[k] = true;
	if (ctx->ref_2d[k]) RB_CUDA(cudaMemsetAsync(ctx->proj_buf[k].as<float2>() + (size_t) mdlX * mdlY, 0, (size_t) mdlX * mdlY * sizeof(float2), ctx->stream));
	return RB_OK;
}

// projector of class k with the full x-pair copy (stage entry points run at arbitrary image sizes)
static RbProjector proj_full(rb_ctx *ctx, int k)
{
	RbProjector p = ctx->proj[k];
	p.mdl2 = ctx->proj2_buf[k].as<float4>();
	p.c2X = p.mdlX; p.c2XY = p.mdlXY; p.c2InitY = p.mdlInitY; p.c2InitZ = p.mdlInitZ;
	p.quad = nullptr;
	return p;
}

// Coarse pass of a pool: point mdl2 at a contiguous x-pair copy of the sphere the coarse window can sample (RbProjector::c2*),
// rebuilt when the reference or the coarse size changed.  RB_COARSE_CORE=0 keeps the full copy (A/B).
static int ensure_coarse_core(rb_ctx *ctx, bool local_search)
{
	static int on = -1, quad_on = -1;
	if (on < 0) { const char *e = getenv("RB_COARSE_CORE"); on = e ? atoi(e) : 1; }
	if (quad_on < 0) { const char *e = getenv("RB_COARSE_QUAD"); quad_on = e ? atoi(e) : 1; }
	const RbModelDev &M = ctx->d_model;
	bool changed = false;
	for (int k = 0; k < M.nr_classes; k++)
	{
		RbProjector &p = ctx->proj[k];
		const int imgMaxR = M.coarse_size / 2;                                           // imgX - 1 (rb_make_projk)
		const int maxR = p.mdlMaxR >= imgMaxR ? imgMaxR : p.mdlMaxR;
		const int R = (int) ceil((double) maxR * p.padding_factor) + 1;                  // |coordinate| <= maxR * pf
		const int cX = R + 1, cY = 2 * R + 2, cInit = -R;
		const bool want = on && !ctx->ref_2d[k] && (size_t) cX * cY * cY * 2 < (size_t) p.mdlXY * p.mdlZ;
		if (!want)
		{
			if (ctx->core_stamp[k] >= 0) { p = proj_full(ctx, k); ctx->core_stamp[k] = -1; changed = true; }
			continue;
		}
		const bool want_quad = quad_on && local_search;          // only the fused local kernel reads the quads
		if (ctx->core_stamp[k] == ctx->ref_version[k] && ctx->core_R[k] == R && (!want_quad || p.quad)) continue;
		RB_CHECK(ctx->proj2c_buf[k].ensure((size_t) cX * cY * cY * sizeof(float4)));
		RB_CHECK(rbk_xpair_core(ctx, p, cX, cY, cInit, cInit, ctx->proj2c_buf[k].as<float4>()));
		p.mdl2 = ctx->proj2c_buf[k].as<float4>();
		p.c2X = cX; p.c2XY = cX * cY; p.c2InitY = cInit; p.c2InitZ = cInit;
		const float4 *old_quad = (ctx->core_stamp[k] == ctx->ref_version[k] && ctx->core_R[k] == R) ? p.quad : nullptr;
		p.quad = old_quad;
		if (want_quad && !old_quad)
		{
			// the xy-quad copy of the same core for the fused local kernel: two 32-byte loads per sample
			RB_CHECK(ctx->proj4c_buf[k].ensure((size_t) cX * cY * cY * 2 * sizeof(float4)));
			RB_CHECK(rbk_xyquad_core(ctx, proj_full(ctx, k), cX, cY, cInit, cInit, ctx->proj4c_buf[k].as<float4>()));
			p.quad = ctx->proj4c_buf[k].as<float4>();
		}
		ctx->core_stamp[k] = ctx->ref_version[k]; ctx->core_R[k] = R;
		changed = true;
	}
	if (changed) RB_CHECK(rb_sync_tables(ctx));
	return RB_OK;
}

extern "C" int rb_set_reference(rb_ctx *ctx, int k, const double *vol, int mdlX, int mdlY, int mdlZ,
                                int initY, int initZ, int maxR, double pf)
{
	const size_t nin = (size_t) mdlX * mdlY * mdlZ;
	RB_CHECK(set_reference_common(ctx, k, mdlX, mdlY, mdlZ, initY, initZ, maxR, pf));
	RB_CHECK(ctx->scratch[2].ensure(nin * 2 * sizeof(double)));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[2].p, vol, nin * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	RB_CHECK(rbk_convert_volume(ctx, ctx->scratch[2].as<double>(), ctx->proj_buf[k].as<float2>(), nin));
	RB_CHECK(rbk_expand_volume(ctx, ctx->proj[k], ctx->proj8_buf[k].as<float4>(), ctx->proj2_buf[k].as<float4>()));
	RB_CHECK(rb_sync_tables(ctx));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->scratch[2].release();
	return RB_OK;
}

extern "C" int rb_set_reference_f32(rb_ctx *ctx, int k, const float *vol, int mdlX, int mdlY, int mdlZ,
                                    int initY, int initZ, int maxR, double pf)
{
	const size_t nin = (size_t) mdlX * mdlY * mdlZ;
	RB_CHECK(set_reference_common(ctx, k, mdlX, mdlY, mdlZ, initY, initZ, maxR, pf));
	RB_CUDA(cudaMemcpyAsync(ctx->proj_buf[k].p, vol, nin * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
	RB_CHECK(rbk_expand_volume(ctx, ctx->proj[k], ctx->proj8_buf[k].as<float4>(), ctx->proj2_buf[k].as<float4>()));
	RB_CHECK(rb_sync_tables(ctx));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

// Projector::computeFourierTransformMap on the device: real-space map -> padded Fourier reference of class k
extern "C" int rb_set_reference_from_map(rb_ctx *ctx, int k, const float *map, int ori_size, int current_size, double padding_factor,
                                         double *power_spectrum)
{
	RB_ARG(ctx && map, "rb_set_reference_from_map: NULL argument");
	RB_ARG(ori_size > 0 && ori_size % 2 == 0 && ori_size / 2 + 1 <= 1024, "rb_set_reference_from_map: box size %d unsupported", ori_size);
	if (current_size <= 0 || current_size > ori_size) current_size = ori_size;
	int padori = (int) floor(padding_factor * ori_size + 0.5);
	padori += padori % 2;
	const double pfe = (double) padori / (double) ori_size;
	const int r_max = std::min(current_size / 2, ori_size / 2);
	const int pad = 2 * ((int) floor(pfe * r_max + 0.5) + 1) + 1;                      // Projector::initialiseData, src/projector.cpp:70
	int mdlX = pad / 2 + 1, mdlY = pad, mdlZ = pad, initY = -((pad - 1) / 2), initZ = initY;
	RB_CHECK(set_reference_common(ctx, k, mdlX, mdlY, mdlZ, initY, initZ, r_max, pfe));
	const size_t n = (size_t) ori_size * ori_size * ori_size;
	RB_CHECK(ctx->scratch[2].ensure(n * sizeof(float)));
	RB_CUDA(cudaMemcpyAsync(ctx->scratch[2].p, map, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	RB_CHECK(rbk_ftmap(ctx, ctx->scratch[2].as<float>(), ori_size, r_max, (float) padding_factor, ctx->proj_buf[k].as<float2>(), pad, power_spectrum));
	RB_CHECK(rbk_expand_volume(ctx, ctx->proj[k], ctx->proj8_buf[k].as<float4>(), ctx->proj2_buf[k].as<float4>()));
	RB_CHECK(rb_sync_tables(ctx));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->scratch[2].release();
	return RB_OK;
}

extern "C" int rb_bp_init(rb_ctx *ctx, int k, int mdlX, int mdlY, int mdlZ, int initY, int initZ, int maxR, double pf)
{
	RB_ARG(ctx, "ctx is NULL");
	RB_ARG(k >= 0 && k < RB_MAX_CLASSES, "class index %d out of range", k);
	RB_ARG(mdlX > 1 && mdlY > 1 && mdlZ >= 1, "rb_bp_init: bad dimensions %dx%dx%d", mdlX, mdlY, mdlZ);
	RB_CUDA(cudaSetDevice(ctx->device));
	ctx->bp_2d[k] = (mdlZ == 1);
	if (mdlZ == 1) { mdlZ = 2; initZ = 0; }             // 2D accumulator: plane 1 only ever receives zeros (fz == 0)
	size_t n = (size_t) mdlX * mdlY * mdlZ;
	RB_CHECK(ctx->bp_buf[k].ensure(n * sizeof(float4)));
	RbBackprojector &b = ctx->bp[k];
	b.vol = ctx->bp_buf[k].as<float4>();
	b.mdlX = mdlX; b.mdlY = mdlY; b.mdlZ = mdlZ; b.mdlInitY = initY; b.mdlInitZ = initZ; b.maxR = maxR; b.padding_factor = (float) pf;
	ctx->has_bp[k] = true;
	RB_CUDA(cudaMemsetAsync(b.vol, 0, n * sizeof(float4), ctx->stream));
	// padded block accumulator of the band-major kernels (3D accumulators; RB_BP_BLOCKS=0: scatter into vol)
	b.blkvol = nullptr; b.blk = nullptr; b.nbx = b.nbxy = 0;
	ctx->bp_blk_dirty[k] = false;
	{
		const char *e = getenv("RB_BP_BLOCKS");
		const int nbx = (mdlX + 3) / 4, nby = (mdlY + 3) / 4, nbz = (mdlZ + 3) / 4;
		const size_t nvox = (size_t) nbx * nby * nbz * 128;
		if (!(e && atoi(e) == 0) && !ctx->bp_2d[k] && nvox < ((size_t) 1 << 31))
		{
			const uint32_t *blk = nullptr;
			RB_CHECK(block_table(ctx, mdlX, mdlY, mdlZ, initY, initZ, &blk));
			RB_CHECK(ctx->bp_blk_buf[k].ensure(nvox * sizeof(float4)));
			b.blkvol = ctx->bp_blk_buf[k].as<float4>(); b.blk = blk; b.nbx = (nbx + 3) / 4; b.nbxy = b.nbx * ((nby + 3) / 4);
			RB_CUDA(cudaMemsetAsync(b.blkvol, 0, nvox * sizeof(float4), ctx->stream));
		}
	}
	RB_CHECK(rb_sync_tables(ctx));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_bp_clear(rb_ctx *ctx, int k)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_bp_clear: accumulator %d not initialised", k);
	const RbBackprojector &b = ctx->bp[k];
	RB_CUDA(cudaMemsetAsync(b.vol, 0, (size_t) b.mdlX * b.mdlY * b.mdlZ * sizeof(float4), ctx->stream));
	if (b.blkvol && ctx->bp_blk_dirty[k])
	{
		RB_CUDA(cudaMemsetAsync(b.blkvol, 0, ctx->bp_blk_buf[k].bytes, ctx->stream));
		ctx->bp_blk_dirty[k] = false;
	}
	return RB_OK;
}

// contributions the band-major kernels left in the padded block accumulator -> vol
int rb_bp_fold(rb_ctx *ctx, int k)
{
	if (!ctx->has_bp[k] || !ctx->bp[k].blkvol || !ctx->bp_blk_dirty[k]) return RB_OK;
	RB_CHECK(rbk_bp_fold(ctx, ctx->bp[k]));
	ctx->bp_blk_dirty[k] = false;
	return RB_OK;
}

extern "C" int rb_bp_get(rb_ctx *ctx, int k, float *real, float *imag, float *weight)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_bp_get: accumulator %d not initialised", k);
	RB_CHECK(rb_bp_fold(ctx, k));
	const RbBackprojector &b = ctx->bp[k];
	size_t n = (size_t) b.mdlX * b.mdlY * b.mdlZ;
	RB_CHECK(ctx->scratch[2].ensure(3 * n * sizeof(float)));
	float *t = ctx->scratch[2].as<float>();
	RB_CHECK(rbk_bp_deinterleave(ctx, b.vol, t, t + n, t + 2 * n, n));
	const size_t nout = ctx->bp_2d[k] ? (size_t) b.mdlX * b.mdlY : n;   // a 2D accumulator hands back its [Y][X] plane
	RB_CUDA(cudaMemcpyAsync(real, t, nout * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(imag, t + n, nout * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(weight, t + 2 * n, nout * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->scratch[2].release();
	return RB_OK;
}

// BackProjector::symmetrise on the device accumulator: Hermitian symmetry of the x = 0 plane + point-group symmetry mates
static int symmetrise_common(rb_ctx *ctx, int k, const double *R, int nsym, int nr_helical_asu, double helical_twist, double helical_rise,
                             int ori_size)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_bp_symmetrise: accumulator %d not initialised", k);
	RB_ARG(!ctx->bp_2d[k], "rb_bp_symmetrise: 2D accumulators are not supported");
	RB_ARG(nsym >= 0 && (nsym == 0 || R), "rb_bp_symmetrise: bad symmetry list");
	RB_ARG(nr_helical_asu < 2 || ori_size > 0, "rb_bp_symmetrise_helical: ori_size must be positive");
	RB_CUDA(cudaSetDevice(ctx->device));
	// all operators in one upload: [nsym][9] point-group matrices, [nhel][9] helical rotations about Z, [nhel] phase ramps
	std::vector<float> r((size_t) std::max(nsym, 0) * 9);
	for (size_t i = 0; i < r.size(); i++) r[i] = (float) R[i];
	int nhel = 0;
	std::vector<float> hz;
	if (nr_helical_asu >= 2)
	{
		// applyHelicalSymmetry (src/backprojector.cpp:2186-2197): hh in [-n/2, n/2 + n%2) without 0, rotation3DMatrix(hh * -twist, 'Z')
		// (src/transformations.cpp:111-117) with setSmallValuesToZero, zshift = hh * rise / (-ori_size * padding_factor) (:2286-2287)
		const int h_min = -nr_helical_asu / 2, h_max = -h_min + nr_helical_asu % 2;
		const double pf = (double) ctx->bp[k].padding_factor;
		for (int hh = h_min; hh < h_max; hh++)
		{
			if (hh == 0) continue;
			const double ang = (double) hh * (-helical_twist) * M_PI / 180.;
			double c = cos(ang), s = sin(ang);
			if (fabs(c) < 1e-6) c = 0.;                                             // XMIPP_EQUAL_ACCURACY (src/macros.h)
			if (fabs(s) < 1e-6) s = 0.;
			const float m[9] = { (float) c, (float) -s, 0.f, (float) s, (float) c, 0.f, 0.f, 0.f, 1.f };
			r.insert(r.end(), m, m + 9);
			hz.push_back(fabs(helical_rise) > 0. ? (float) ((double) hh * helical_rise / (-(double) ori_size * pf)) : 0.f);
			nhel++;
		}
		r.insert(r.end(), hz.begin(), hz.end());
	}
	float *d_R = nullptr, *d_hR = nullptr, *d_hz = nullptr;
	if (!r.empty())
	{
		RB_CHECK(upload(ctx, ctx->scratch[3], r.data(), r.size() * 4));
		d_R = ctx->scratch[3].as<float>();
		d_hR = d_R + (size_t) nsym * 9;
		d_hz = d_hR + (size_t) nhel * 9;
	}
	RB_CHECK(rb_bp_fold(ctx, k));
	RB_CHECK(rbk_bp_symmetrise(ctx, ctx->bp[k], ctx->recon_buf[1], d_R, nsym, d_hR, d_hz, nhel));
	return RB_OK;
}

extern "C" int rb_bp_symmetrise(rb_ctx *ctx, int k, const double *R, int nsym)
{
	return symmetrise_common(ctx, k, R, nsym, 1, 0., 0., 0);
}

extern "C" int rb_bp_symmetrise_helical(rb_ctx *ctx, int k, const double *R, int nsym, int nr_helical_asu, double helical_twist,
                                        double helical_rise, int ori_size)
{
	return symmetrise_common(ctx, k, R, nsym, nr_helical_asu, helical_twist, helical_rise, ori_size);
}

// BackProjector::reconstruct (default skip_gridding branch) on the device
static int reconstruct_common(rb_ctx *ctx, int k, int ori_size, const double *tau2, int n_tau2, double tau2_fudge, int minres_map,
                              int max_iter_preweight, double normalise, float *vol_out)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_reconstruct: accumulator %d not initialised", k);
	RB_ARG(!ctx->bp_2d[k], "rb_reconstruct: 2D accumulators are not supported yet");
	RB_ARG(vol_out && ori_size > 0 && ori_size % 2 == 0, "rb_reconstruct: bad arguments");
	RB_ARG(!tau2 || n_tau2 > 0, "rb_reconstruct: tau2 without a length");
	RB_CUDA(cudaSetDevice(ctx->device));
	const size_t n = (size_t) ori_size * ori_size * ori_size;
	RB_CHECK(ctx->scratch[2].ensure(n * sizeof(float)));
	double *d_tau2 = nullptr;
	if (tau2)
	{
		RB_CHECK(ctx->scratch[3].ensure((size_t) n_tau2 * 8));
		RB_CUDA(cudaMemcpyAsync(ctx->scratch[3].p, tau2, (size_t) n_tau2 * 8, cudaMemcpyHostToDevice, ctx->stream));
		d_tau2 = ctx->scratch[3].as<double>();
	}
	RB_CHECK(rb_bp_fold(ctx, k));
	RB_CHECK(rbk_reconstruct(ctx, ctx->bp[k], ori_size, d_tau2, n_tau2, tau2_fudge, minres_map, ctx->scratch[2].as<float>(), max_iter_preweight, normalise));
	RB_CUDA(cudaMemcpyAsync(vol_out, ctx->scratch[2].p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->scratch[2].release();
	return RB_OK;
}

extern "C" int rb_reconstruct(rb_ctx *ctx, int k, int ori_size, const double *tau2, int n_tau2, double tau2_fudge, int minres_map, float *vol_out)
{
	return reconstruct_common(ctx, k, ori_size, tau2, n_tau2, tau2_fudge, minres_map, 0, 1., vol_out);
}

extern "C" int rb_reconstruct_gridding(rb_ctx *ctx, int k, int ori_size, const double *tau2, int n_tau2, double tau2_fudge, int minres_map,
                                       int max_iter_preweight, double normalise, float *vol_out)
{
	RB_ARG(max_iter_preweight > 0 && max_iter_preweight <= 100, "rb_reconstruct_gridding: max_iter_preweight %d out of range", max_iter_preweight);
	return reconstruct_common(ctx, k, ori_size, tau2, n_tau2, tau2_fudge, minres_map, max_iter_preweight, normalise, vol_out);
}

// BackProjector::updateSSNRarrays on the device accumulator
extern "C" int rb_update_ssnr(rb_ctx *ctx, int k, int ori_size, double tau2_fudge, double *tau2_io, double *sigma2_out,
                              double *data_vs_prior_out, double *fourier_coverage_out, const double *fsc, const double *avgctf2,
                              int update_tau2_with_fsc, int is_whole_instead_of_half)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_update_ssnr: accumulator %d not initialised", k);
	RB_ARG(tau2_io && sigma2_out && data_vs_prior_out && fourier_coverage_out && ori_size > 0, "rb_update_ssnr: NULL spectrum");
	RB_CUDA(cudaSetDevice(ctx->device));
	RB_CHECK(rb_bp_fold(ctx, k));
	return rbk_update_ssnr(ctx, ctx->bp[k], ctx->bp_2d[k], ori_size, tau2_fudge, tau2_io, sigma2_out, data_vs_prior_out, fourier_coverage_out,
	                       fsc, avgctf2, update_tau2_with_fsc != 0, is_whole_instead_of_half != 0);
}

extern "C" int rb_bp_device_buffer(rb_ctx *ctx, int k, void **dptr, size_t *n_floats)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_bp_device_buffer: accumulator %d not initialised", k);
	RB_CUDA(cudaSetDevice(ctx->device));
	RB_CHECK(rb_bp_fold(ctx, k));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));      // the caller works on the buffer from its own stream
	const RbBackprojector &b = ctx->bp[k];
	if (dptr) *dptr = b.vol;
	if (n_floats) *n_floats = (size_t) b.mdlX * b.mdlY * b.mdlZ * 4;
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// sampling
// ---------------------------------------------------------------------------------------------
extern "C" int rb_set_sampling(rb_ctx *ctx, const rb_sampling *s)
{
	RB_ARG(ctx && s, "rb_set_sampling: NULL argument");
	RB_ARG(s->n_dir > 0 && s->n_psi > 0 && s->n_trans > 0, "rb_set_sampling: empty sampling");
	RB_ARG(s->rot && s->tilt && s->psi && s->trans_x && s->trans_y, "rb_set_sampling: NULL table");
	RB_ARG(s->n_over_rot >= 1 && s->n_over_trans >= 1, "rb_set_sampling: oversampling factors must be >= 1");
	RB_ARG(s->n_over_rot == 1 || (s->over_rot && s->over_tilt && s->over_psi), "rb_set_sampling: oversampled orientations missing");
	ctx->samp_version++;
	RB_ARG(s->n_over_trans == 1 || (s->over_trans_x && s->over_trans_y), "rb_set_sampling: oversampled translations missing");
	RB_ARG(ctx->has_model, "rb_set_sampling: call rb_set_model first (translations are scaled by ori_size)");
	if ((long long) s->n_trans * s->n_over_trans > 2048)
	{
		rb_set_error("rb_set_sampling: %d translations x %d oversampling exceeds the supported 2048 (ERR_TRANSLIM)", s->n_trans, s->n_over_trans);
		return RB_ERR_TRANSLIM;
	}
	RB_CUDA(cudaSetDevice(ctx->device));
	ctx->h_samp = *s;
	const size_t no = (size_t) s->n_dir * s->n_psi;
	const int T = s->n_trans, Tf = T * s->n_over_trans;
	// angles (fp64 for the fine pass, fp32 for the coarse kernel like AccProjectorPlan's XFLOAT alphas/betas/gammas)
	RB_CHECK(upload(ctx, ctx->s_rot, s->rot, s->n_dir * sizeof(double)));
	RB_CHECK(upload(ctx, ctx->s_tilt, s->tilt, s->n_dir * sizeof(double)));
	RB_CHECK(upload(ctx, ctx->s_psi, s->psi, s->n_psi * sizeof(double)));
	RB_CHECK(ctx->s_coarse_eulers.ensure(no * 9 * sizeof(float)));
	ctx->coarse_lr = RbLR();
	RB_CHECK(rbk_make_coarse_eulers(ctx, ctx->s_rot.as<double>(), ctx->s_tilt.as<double>(), ctx->s_psi.as<double>(),
	                                s->n_dir, s->n_psi, ctx->coarse_lr, ctx->s_coarse_eulers.as<float>()));
	if (s->n_over_rot > 1 || s->over_rot)
	{
		RB_CHECK(upload(ctx, ctx->s_over_rot, s->over_rot, no * s->n_over_rot * sizeof(double)));
		RB_CHECK(upload(ctx, ctx->s_over_tilt, s->over_tilt, no * s->n_over_rot * sizeof(double)));
		RB_CHECK(upload(ctx, ctx->s_over_psi, s->over_psi, no * s->n_over_rot * sizeof(double)));
	}
	// translations: pixels (for priors) and -2*pi*shift/ori_size (acc_ml_optimiser_impl.h:1239-1241, 1549-1551)
	const double os = (double) ctx->h_model.ori_size;
	std::vector<float> ctxv(T), ctyv(T), ftxv(Tf), ftyv(Tf);
	std::vector<double> otx(Tf), oty(Tf);
	for (int t = 0; t < T; t++)
	{
		ctxv[t] = (float) (-2 * M_PI * s->trans_x[t] / os);
		ctyv[t] = (float) (-2 * M_PI * s->trans_y[t] / os);
	}
	for (int t = 0; t < Tf; t++)
	{
		otx[t] = s->over_trans_x ? s->over_trans_x[t] : s->trans_x[t];
		oty[t] = s->over_trans_y ? s->over_trans_y[t] : s->trans_y[t];
		ftxv[t] = (float) (-2 * M_PI * otx[t] / os);
		ftyv[t] = (float) (-2 * M_PI * oty[t] / os);
	}
	// band-major diff2 pass: do the oversampled translations factorise into coarse translation + fixed offset?
	{
		const int NOT = s->n_over_trans;
		bool sep = true;
		for (int t = 0; t < T && sep; t++)
			for (int j = 0; j < NOT; j++)
			{
				const double dx = otx[t * NOT + j] - s->trans_x[t], dy = oty[t * NOT + j] - s->trans_y[t];
				if (fabs(dx - (otx[j] - s->trans_x[0])) > 1e-9 || fabs(dy - (oty[j] - s->trans_y[0])) > 1e-9) { sep = false; break; }
			}
		ctx->band_separable = sep;
		ctx->h_band_u.assign((size_t) 2 * (T + NOT), 0.);
		for (int t = 0; t < T; t++) { ctx->h_band_u[t] = -s->trans_x[t] / os; ctx->h_band_u[T + t] = -s->trans_y[t] / os; }
		for (int j = 0; j < NOT; j++) { ctx->h_band_u[2 * T + j] = -(otx[j] - s->trans_x[0]) / os; ctx->h_band_u[2 * T + NOT + j] = -(oty[j] - s->trans_y[0]) / os; }
	}
	RB_CHECK(upload(ctx, ctx->s_ctx, ctxv.data(), T * 4)); RB_CHECK(upload(ctx, ctx->s_cty, ctyv.data(), T * 4));
	RB_CHECK(upload(ctx, ctx->s_ftx, ftxv.data(), Tf * 4)); RB_CHECK(upload(ctx, ctx->s_fty, ftyv.data(), Tf * 4));
	RB_CHECK(upload(ctx, ctx->s_tx, s->trans_x, T * 8)); RB_CHECK(upload(ctx, ctx->s_ty, s->trans_y, T * 8));
	RB_CHECK(upload(ctx, ctx->s_otx, otx.data(), Tf * 8)); RB_CHECK(upload(ctx, ctx->s_oty, oty.data(), Tf * 8));

	RbSamplingDev &d = ctx->d_samp;
	d.n_dir = s->n_dir; d.n_psi = s->n_psi; d.n_over_rot = s->n_over_rot; d.n_trans = T; d.n_over_trans = s->n_over_trans;
	d.coarse_eulers = ctx->s_coarse_eulers.as<float>();
	const bool have_over = (s->over_rot != nullptr);
	d.over_rot = have_over ? ctx->s_over_rot.as<double>() : nullptr;
	d.over_tilt = have_over ? ctx->s_over_tilt.as<double>() : nullptr;
	d.over_psi = have_over ? ctx->s_over_psi.as<double>() : nullptr;
	d.rot = ctx->s_rot.as<double>(); d.tilt = ctx->s_tilt.as<double>(); d.psi = ctx->s_psi.as<double>();
	d.ctx = ctx->s_ctx.as<float>(); d.cty = ctx->s_cty.as<float>(); d.ftx = ctx->s_ftx.as<float>(); d.fty = ctx->s_fty.as<float>();
	d.trans_x = ctx->s_tx.as<double>(); d.trans_y = ctx->s_ty.as<double>();
	d.over_trans_x = ctx->s_otx.as<double>(); d.over_trans_y = ctx->s_oty.as<double>();
	RB_CUDA(cudaStreamSynchronize(ctx->stream));   // host vectors above are temporaries
	ctx->h_samp.rot = ctx->h_samp.tilt = ctx->h_samp.psi = nullptr;
	ctx->has_sampling = true;
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------
static inline int iround(double x) { return (int) (x > 0 ? floor(x + 0.5) : -floor(-x + 0.5)); }

// pixels with Mresol >= 0 for window n (src/ml_optimiser.cpp:5784-5811), in FFTW order
// full_x0: keep the redundant half of the x = 0 column (jp == 0, ip < 0).  Mresol excludes it (Minvsigma2 is zero there),
// but the cross-correlation kernels weight every pixel with 1 / sqrtXi2^2 (buildCorrImage), so it contributes there.
// dead > 0: drop the rows the reference's fine / wavg kernels skip when the references end inside the window
// (maxR = dead < n / 2: rows |ip| > maxR keep only their pixel jp == maxR, cpu_kernels/diff2.h:347-355)
// all_pixels: the whole window, corners included (shell index clamped) — the cross-correlation kernels weight EVERY pixel, and a
// contracting MBL brings corner pixels inside the reference (elsewhere they project to zero and contribute nothing)
static void make_pixlist(int n, std::vector<uint32_t> &out, bool full_x0 = false, int dead = 0, bool all_pixels = false)
{
	const int xs = n / 2 + 1;
	out.clear();
	for (int iy = 0; iy < n; iy++)
	{
		const int ip = iy < xs ? iy : iy - n;
		for (int jp = 0; jp < xs; jp++)
		{
			const int ires = iround(sqrt((double) (ip * ip + jp * jp)));
			if (dead > 0 && abs(ip) > dead && jp != dead) continue;
			if ((ires < xs || all_pixels) && (full_x0 || !(jp == 0 && ip < 0))) out.push_back(rb_pack_pix(jp, ip, std::min(ires, xs - 1)));
		}
	}
}

// the same pixel set as row runs + a dense shell map
static void make_rows(int n, std::vector<RbRow> &rows, std::vector<short> &ires_map, bool full_x0 = false, int dead = 0, bool all_pixels = false)
{
	const int xs = n / 2 + 1;
	rows.clear();
	ires_map.assign((size_t) n * xs, -1);
	for (int iy = 0; iy < n; iy++)
	{
		const int ip = iy < xs ? iy : iy - n;
		int lo = -1, hi = -1;
		for (int jp = 0; jp < xs; jp++)
		{
			const int ires = iround(sqrt((double) (ip * ip + jp * jp)));
			if (dead > 0 && abs(ip) > dead && jp != dead) continue;
			if ((ires < xs || all_pixels) && (full_x0 || !(jp == 0 && ip < 0)))
			{
				ires_map[(size_t) iy * xs + jp] = (short) std::min(ires, xs - 1);
				if (lo < 0) lo = jp;
				hi = jp;
			}
		}
		if (lo >= 0) rows.push_back(RbRow{(short) iy, (short) ip, (short) lo, (short) hi});
	}
}

extern "C" int rb_set_model(rb_ctx *ctx, const rb_model *m)
{
	RB_ARG(ctx && m, "rb_set_model: NULL argument");
	RB_ARG(m->nr_classes >= 1 && m->nr_classes <= RB_MAX_CLASSES, "rb_set_model: nr_classes %d unsupported", m->nr_classes);
	RB_ARG(m->ori_size > 0 && m->ori_size % 2 == 0 && m->ori_size <= 1000, "rb_set_model: ori_size %d unsupported (even, <= 1000)", m->ori_size);
	RB_ARG(m->current_size > 0 && m->current_size % 2 == 0 && m->current_size <= m->ori_size, "rb_set_model: bad current_size %d", m->current_size);
	RB_ARG(m->coarse_size > 0 && m->coarse_size % 2 == 0 && m->coarse_size <= m->current_size, "rb_set_model: bad coarse_size %d", m->coarse_size);
	RB_ARG(m->sigma2_noise && m->pdf_class && m->nr_optics_groups >= 1 && m->nr_groups >= 1, "rb_set_model: NULL table");
	RB_ARG(!m->do_scale_correction || m->scale_correction, "rb_set_model: scale_correction missing");
	RB_ARG(m->sigma2_fudge > 0., "rb_set_model: sigma2_fudge must be > 0");
	RB_CUDA(cudaSetDevice(ctx->device));
	ctx->h_model = *m;
	ctx->model_version++;
	const int nshell = m->ori_size / 2 + 1, K = m->nr_classes;
	RB_ARG(m->ref_max_r >= 0, "rb_set_model: bad ref_max_r %d", m->ref_max_r);
	const int dead = (m->ref_max_r > 0 && m->ref_max_r < m->current_size / 2) ? m->ref_max_r : 0;
	ctx->fine_dead_maxR = dead;
	std::vector<uint32_t> pc, pf;
	make_pixlist(m->coarse_size, pc); make_pixlist(m->current_size, pf, false, dead);
	{
		// The coarse kernels are order-agnostic over this list, and the fused kernel projects 16 consecutive entries per
		// warp-row: grouping them as W x H image tiles instead of row segments makes neighbouring lanes land in
		// neighbouring voxels, i.e. fewer distinct 128-byte lines per divergent load (RB_PIX_TILE=WxH, 16x1 = row order).
		auto tile_sort = [](std::vector<uint32_t> &v, int tw, int th) {
			if (th <= 1) return;
			std::stable_sort(v.begin(), v.end(), [tw, th](uint32_t a, uint32_t b) {
				const int ya = rb_pix_y(a) + 512, yb = rb_pix_y(b) + 512, xa = rb_pix_x(a), xb = rb_pix_x(b);
				if (ya / th != yb / th) return ya / th < yb / th;
				if (xa / tw != xb / tw) return xa / tw < xb / tw;
				if (ya != yb) return ya < yb;
				return xa < xb;
			});
		};
		int tw = 4, th = 4;                          // measured (tools/sweep_tiles.sh): 16x1 8.94 ms, 4x4 7.88 ms, 2x4 7.96 ms, 8x2 8.52 ms
		if (const char *e = getenv("RB_PIX_TILE")) { if (sscanf(e, "%dx%d", &tw, &th) != 2 || tw < 1 || th < 1) { tw = 4; th = 4; } }
		tile_sort(pc, tw, th);
		// the store stage walks the fine list with two lanes per pixel: tiles put a warp's reductions into one small 3D patch
		int fw = 4, fh = 4;                          // store stage: 16x1 5.50 ms, 4x4 5.32 ms, 4x2 5.28 ms
		if (const char *e = getenv("RB_PIX_TILE_F")) { if (sscanf(e, "%dx%d", &fw, &fh) != 2 || fw < 1 || fh < 1) { fw = 4; fh = 4; } }
		tile_sort(pf, fw, fh);
	}
	RB_CHECK(upload(ctx, ctx->m_pix_c, pc.data(), pc.size() * 4));
	RB_CHECK(upload(ctx, ctx->m_pix_f, pf.data(), pf.size() * 4));
	// --no_map: Minvsigma2 is one on EVERY pixel in the back-projection (acc_ml_optimiser_impl.h:3110-3115), so the redundant
	// half of the x = 0 column, which Mresol excludes, is back-projected as well: the store stage then walks the full list
	// ... and the reference's back-projection kernel has no row rule (BP.h:559-565): with dead rows the store list keeps them
	size_t n_store = pf.size();
	if (!m->do_map || dead > 0)
	{
		std::vector<uint32_t> pfull;
		make_pixlist(m->current_size, pfull, !m->do_map, 0);
		n_store = pfull.size();
		RB_CHECK(upload(ctx, ctx->m_cc[5], pfull.data(), pfull.size() * 4));
	}
	// band-major kernels: the store-stage pixel set sorted by |r| (ties by angle), the diff2 set first
	std::vector<uint32_t> prs;
	{
		std::vector<uint32_t> all;
		make_pixlist(m->current_size, all, !m->do_map, 0);
		auto in_d2 = [dead](uint32_t v) {
			return !(rb_pix_x(v) == 0 && rb_pix_y(v) < 0) && !(dead > 0 && abs(rb_pix_y(v)) > dead && rb_pix_x(v) != dead);
		};
		auto radial = [](uint32_t a, uint32_t b) {
			const int xa = rb_pix_x(a), ya = rb_pix_y(a), xb = rb_pix_x(b), yb = rb_pix_y(b);
			const int ra = xa * xa + ya * ya, rb2 = xb * xb + yb * yb;
			if (ra != rb2) return ra < rb2;
			const double ta = atan2((double) ya, (double) xa), tb = atan2((double) yb, (double) xb);
			if (ta != tb) return ta < tb;
			return a < b;
		};
		std::vector<uint32_t> extra;
		for (uint32_t v : all) (in_d2(v) ? prs : extra).push_back(v);
		std::sort(prs.begin(), prs.end(), radial); std::sort(extra.begin(), extra.end(), radial);
		const int nd2 = (int) prs.size();
		prs.insert(prs.end(), extra.begin(), extra.end());
		const int nst = (int) prs.size();
		const int npad = (nst + 127) / 128 * 128;
		prs.resize(npad, rb_pack_pix(0, 0, 0));
		RB_CHECK(upload(ctx, ctx->m_pix_rs, prs.data(), prs.size() * 4));
		ctx->d_model.pix_rs = ctx->m_pix_rs.as<uint32_t>();
		ctx->d_model.nv_rs_d2 = nd2; ctx->d_model.nv_rs_st = nst; ctx->d_model.nv_rs_pad = npad;
	}
	std::vector<RbRow> rc, rf;
	std::vector<short> ic, iff;
	make_rows(m->coarse_size, rc, ic); make_rows(m->current_size, rf, iff, false, dead);
	RB_CHECK(upload(ctx, ctx->m_rows_c, rc.data(), rc.size() * sizeof(RbRow)));
	RB_CHECK(upload(ctx, ctx->m_rows_f, rf.data(), rf.size() * sizeof(RbRow)));
	RB_CHECK(upload(ctx, ctx->m_ires_c, ic.data(), ic.size() * sizeof(short)));
	RB_CHECK(upload(ctx, ctx->m_ires_f, iff.data(), iff.size() * sizeof(short)));
	// Minvsigma2 per shell (src/ml_optimiser.cpp:6868-6879); entry 0 holds the DC value restored for the
	// store stage (acc_ml_optimiser_impl.h:2586), the diff2 kernels ignore it
	std::vector<float> mv((size_t) m->nr_optics_groups * nshell);
	for (int g = 0; g < m->nr_optics_groups; g++)
		for (int i = 0; i < nshell; i++)
			mv[(size_t) g * nshell + i] = (float) (1. / (m->sigma2_fudge * m->sigma2_noise[(size_t) g * nshell + i]));
	RB_CHECK(upload(ctx, ctx->m_minvs2, mv.data(), mv.size() * 4));
	{
		// [K] pdf_class, then [K][2] prior_offset_class
		std::vector<double> pc((size_t) 3 * K, 0.);
		for (int k = 0; k < K; k++) pc[k] = m->pdf_class[k];
		if (m->prior_offset_class) for (int k = 0; k < 2 * K; k++) pc[K + k] = m->prior_offset_class[k];
		RB_CHECK(upload(ctx, ctx->m_pdf_class, pc.data(), pc.size() * sizeof(double)));
	}
	std::vector<unsigned char> dvp((size_t) K * nshell, 0);
	if (m->data_vs_prior_class)
		for (size_t i = 0; i < dvp.size(); i++) dvp[i] = m->data_vs_prior_class[i] > 3.;
	RB_CHECK(upload(ctx, ctx->m_dvp, dvp.data(), dvp.size()));
	ctx->h_scale_correction.assign(m->nr_groups, 1.);
	if (m->scale_correction) ctx->h_scale_correction.assign(m->scale_correction, m->scale_correction + m->nr_groups);

	RbModelDev &d = ctx->d_model;
	d.nr_classes = K; d.ori_size = m->ori_size; d.coarse_size = m->coarse_size; d.current_size = m->current_size; d.nshell = nshell;
	d.Npc = m->coarse_size * (m->coarse_size / 2 + 1); d.Npf = m->current_size * (m->current_size / 2 + 1);
	d.nvc = (int) pc.size(); d.nvf = (int) pf.size();
	d.pix_c = ctx->m_pix_c.as<uint32_t>(); d.pix_f = ctx->m_pix_f.as<uint32_t>();
	d.pix_store = (m->do_map && dead == 0) ? d.pix_f : ctx->m_cc[5].as<uint32_t>(); d.nv_store = (int) n_store;
	d.dead_maxR = dead;
	d.nrows_c = (int) rc.size(); d.nrows_f = (int) rf.size();
	d.rows_c = ctx->m_rows_c.as<RbRow>(); d.rows_f = ctx->m_rows_f.as<RbRow>();
	d.ires_c = ctx->m_ires_c.as<short>(); d.ires_f = ctx->m_ires_f.as<short>();
	// pixel sets of the two diff2 passes: the Mresol sets above, or with the cross-correlation criterion every pixel
	// of the window (corners included: they project to zero unless a contracting MBL brings them inside the reference)
	d.d2_nvc = d.nvc; d.d2_pix_c = d.pix_c;
	d.d2_nrows_c = d.nrows_c; d.d2_rows_c = d.rows_c; d.d2_ires_c = d.ires_c;
	d.d2_nrows_f = d.nrows_f; d.d2_rows_f = d.rows_f; d.d2_ires_f = d.ires_f;
	if (m->do_cc)
	{
		std::vector<uint32_t> pcc;
		std::vector<RbRow> rcc, rfc;
		std::vector<short> icc, ifc;
		// unlike the Gaussian coarse kernel, the cross-correlation coarse kernel has the row rule too (diff2.h:657-666)
		const int dead_c = (m->ref_max_r > 0 && m->ref_max_r < m->coarse_size / 2) ? m->ref_max_r : 0;
		make_pixlist(m->coarse_size, pcc, true, dead_c, true);
		make_rows(m->coarse_size, rcc, icc, true, dead_c, true); make_rows(m->current_size, rfc, ifc, true, dead, true);
		RB_CHECK(upload(ctx, ctx->m_cc[0], pcc.data(), pcc.size() * 4));
		RB_CHECK(upload(ctx, ctx->m_cc[1], rcc.data(), rcc.size() * sizeof(RbRow)));
		RB_CHECK(upload(ctx, ctx->m_cc[2], icc.data(), icc.size() * sizeof(short)));
		RB_CHECK(upload(ctx, ctx->m_cc[3], rfc.data(), rfc.size() * sizeof(RbRow)));
		RB_CHECK(upload(ctx, ctx->m_cc[4], ifc.data(), ifc.size() * sizeof(short)));
		d.d2_nvc = (int) pcc.size(); d.d2_pix_c = ctx->m_cc[0].as<uint32_t>();
		d.d2_nrows_c = (int) rcc.size(); d.d2_rows_c = ctx->m_cc[1].as<RbRow>(); d.d2_ires_c = ctx->m_cc[2].as<short>();
		d.d2_nrows_f = (int) rfc.size(); d.d2_rows_f = ctx->m_cc[3].as<RbRow>(); d.d2_ires_f = ctx->m_cc[4].as<short>();
	}
	d.minvs2 = ctx->m_minvs2.as<float>();
	d.pdf_class = ctx->m_pdf_class.as<double>();
	d.prior_offset_class = m->prior_offset_class ? ctx->m_pdf_class.as<double>() + K : nullptr;
	d.pdf_direction = nullptr;
	d.dvp_gt3 = ctx->m_dvp.as<unsigned char>();
	d.pixel_size = m->pixel_size;
	d.s2off = (m->offset_range > 0.) ? (m->offset_range * m->offset_range) / 9. : m->sigma2_offset;   // :1916-1926
	d.adaptive_fraction = m->adaptive_fraction; d.maximum_significants = m->maximum_significants;
	d.do_ctf_correction = m->do_ctf_correction; d.refs_are_ctf_corrected = m->refs_are_ctf_corrected;
	d.do_scale_correction = m->do_scale_correction; d.do_map = m->do_map; d.ctf_premultiplied = m->ctf_premultiplied;
	d.bp_circle_bound = m->bp_circle_bound;
	d.do_cc = m->do_cc;
	d.do_grad = m->do_grad;
	d.do_skip_rotate = m->do_skip_rotate;
	// pdf_direction needs n_dir, which belongs to the sampling: keep a host copy until both are known
	ctx->m_pdf_dir.release();
	if (m->pdf_direction && ctx->has_sampling)
	{
		RB_CHECK(upload(ctx, ctx->m_pdf_dir, m->pdf_direction, (size_t) K * ctx->d_samp.n_dir * sizeof(double)));
		d.pdf_direction = ctx->m_pdf_dir.as<double>();
	}
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->h_sigma2_noise.assign(m->sigma2_noise, m->sigma2_noise + (size_t) m->nr_optics_groups * nshell);
	ctx->h_model.sigma2_noise = nullptr; ctx->h_model.scale_correction = nullptr; ctx->h_model.pdf_class = nullptr;
	ctx->h_model.data_vs_prior_class = nullptr; ctx->h_model.prior_offset_class = nullptr;
	ctx->has_model = true;
	return RB_OK;
}

// pdf_direction depends on both model (values) and sampling (n_dir): dedicated setter so call order is free
extern "C" int rb_set_pdf_direction(rb_ctx *ctx, const double *pdf_direction)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && pdf_direction && ctx->has_model && ctx->has_sampling, "rb_set_pdf_direction: needs model and sampling");
	RB_CHECK(upload(ctx, ctx->m_pdf_dir, pdf_direction, (size_t) ctx->d_model.nr_classes * ctx->d_samp.n_dir * sizeof(double)));
	ctx->d_model.pdf_direction = ctx->m_pdf_dir.as<double>();
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// pool upload
// ---------------------------------------------------------------------------------------------
static size_t env_size(const char *name, size_t dflt)
{
	const char *v = getenv(name);
	if (!v || !*v) return dflt;
	return (size_t) strtoull(v, nullptr, 10);
}

// Shared by rb_pool_upload (prepared Fourier images from the host) and rb_pool_prepare (raw images, prepared on the device):
// per-particle metadata, prior lists, work buffers; copy_images == false leaves Fimg / Fimg_nomask / Fctf to the caller.
static int pool_setup(rb_ctx *ctx, int slot, const rb_particles *pool, bool copy_images)
{
	RB_ARG(ctx && pool, "rb_pool_upload: NULL argument");
	RB_ARG(slot >= 0 && slot < RB_NUM_SLOTS, "rb_pool_upload: slot %d out of range", slot);
	if (!ctx->has_model || !ctx->has_sampling) { rb_set_error("rb_pool_upload: set model and sampling first"); return RB_ERR_STATE; }
	const int P = pool->n_particles;
	RB_ARG(P > 0, "rb_pool_upload: empty pool");
	RB_ARG(pool->group_id && pool->optics_group && pool->highres_Xi2 && pool->old_offset && pool->prior_offset,
	       "rb_pool_upload: NULL particle array");
	RB_ARG(!copy_images || (pool->Fimg && pool->Fimg_nomask), "rb_pool_upload: NULL image array");
	RB_ARG(!copy_images || !ctx->h_model.do_ctf_correction || pool->Fctf, "rb_pool_upload: Fctf missing with do_ctf_correction");
	const bool priors = pool->dir_idx != nullptr;
	RB_ARG(!priors || (pool->dir_off && pool->dir_prior && pool->psi_off && pool->psi_idx && pool->psi_prior), "rb_pool_upload: incomplete prior lists");
	RB_ARG(priors || ctx->d_model.pdf_direction, "rb_pool_upload: pdf_direction needed without orientational priors");
	RB_CUDA(cudaSetDevice(ctx->device));
	PoolSlot &s = ctx->slot[slot];
	const RbModelDev &M = ctx->d_model;
	const RbSamplingDev &S = ctx->d_samp;
	const int K = M.nr_classes, T = S.n_trans;

	s.P = P; s.has_priors = priors;
	s.h_meta.resize(P);
	long long coff = 0, poff = 0; int max_no = 0;
	s.max_bp_off = 0;
	s.lr = RbLR();
	if (pool->mat_left || pool->mat_right)
	{
		RB_ARG(!ctx->ref_2d[0], "rb_pool_upload: mat_left / mat_right need 3D references (make_eulers_2D takes neither)");
		if (pool->mat_left) { s.lr.doL = 1; for (int i = 0; i < 9; i++) s.lr.L[i] = pool->mat_left[i]; }
		if (pool->mat_right) { s.lr.doR = 1; for (int i = 0; i < 9; i++) s.lr.R[i] = pool->mat_right[i]; }
	}
	for (int p = 0; p < P; p++)
	{
		RbPartMeta &m = s.h_meta[p];
		if (priors)
		{
			m.nd = pool->dir_off[p + 1] - pool->dir_off[p]; m.np = pool->psi_off[p + 1] - pool->psi_off[p];
			m.dir_off = pool->dir_off[p]; m.psi_off = pool->psi_off[p];
			RB_ARG(m.nd > 0 && m.np > 0, "rb_pool_upload: particle %d has an empty orientation list", p);
		}
		else { m.nd = S.n_dir; m.np = S.n_psi; m.dir_off = -1; m.psi_off = -1; }
		m.group = pool->group_id[p]; m.og = pool->optics_group[p];
		RB_ARG(m.group >= 0 && m.group < ctx->h_model.nr_groups && m.og >= 0 && m.og < ctx->h_model.nr_optics_groups,
		       "rb_pool_upload: particle %d group/optics group out of range", p);
		double sc = ctx->h_model.do_scale_correction ? ctx->h_scale_correction[m.group] : 1.;
		m.scale = (float) sc;
		float ps = 1.f;
		if (ctx->h_model.do_scale_correction)
		{
			ps = (float) sc;
			if (ps > 10000.f) { rb_set_error("rlnMicrographScaleCorrection %g of group %d is too high (ERRHIGHSCALE)", sc, m.group + 1); return RB_ERR_ARG; }
			if (ps < 0.001f) ps = 0.001f;                                              // :3069-3078
		}
		m.part_scale = ps;
		m.xi2_half = (float) (pool->highres_Xi2[p] / 2.);
		m.bp_off = pool->bp_offset ? pool->bp_offset[p] : 0;
		RB_ARG(m.bp_off >= 0 && m.bp_off + K <= RB_MAX_CLASSES, "rb_pool_upload: particle %d: accumulator offset %d out of range", p, m.bp_off);
		s.max_bp_off = std::max(s.max_bp_off, m.bp_off);
		m.oldx = pool->old_offset[2 * p]; m.oldy = pool->old_offset[2 * p + 1];
		if (pool->pre_shift) { m.oldx += pool->pre_shift[2 * p]; m.oldy += pool->pre_shift[2 * p + 1]; }   // part of every sampled offset
		m.prx = pool->prior_offset[2 * p]; m.pry = pool->prior_offset[2 * p + 1];
		m.coarse_off = coff; m.prior_off = poff;
		const long long no = (long long) m.nd * m.np;
		coff += (long long) K * no * T; poff += (long long) K * no;
		max_no = std::max<long long>(max_no, no);
	}
	s.total_coarse = coff; s.total_prior = poff; s.max_no = max_no;

	cudaStream_t cs = ctx->copy_stream;
	const size_t img_bytes = (size_t) P * M.Npf * sizeof(float2);
	RB_CHECK(s.Fimg.ensure(img_bytes)); RB_CHECK(s.Fnomask.ensure(img_bytes));
	if (ctx->h_model.do_ctf_correction) RB_CHECK(s.Fctf.ensure(img_bytes / 2));
	if (copy_images)
	{
		RB_CUDA(cudaMemcpyAsync(s.Fimg.p, pool->Fimg, img_bytes, cudaMemcpyHostToDevice, cs));
		RB_CUDA(cudaMemcpyAsync(s.Fnomask.p, pool->Fimg_nomask, img_bytes, cudaMemcpyHostToDevice, cs));
		if (ctx->h_model.do_ctf_correction) RB_CUDA(cudaMemcpyAsync(s.Fctf.p, pool->Fctf, img_bytes / 2, cudaMemcpyHostToDevice, cs));
	}
	RB_CHECK(s.meta.ensure(P * sizeof(RbPartMeta)));
	RB_CUDA(cudaMemcpyAsync(s.meta.p, s.h_meta.data(), P * sizeof(RbPartMeta), cudaMemcpyHostToDevice, cs));
	s.has_order = false;
	if (priors && env_size("RB_COARSE_ORDER", 1) != 0)
	{
		s.h_order.resize(P);
		for (int p = 0; p < P; p++) s.h_order[p] = p;
		const int *doff = pool->dir_off, *didx = pool->dir_idx;
		std::stable_sort(s.h_order.begin(), s.h_order.end(), [doff, didx](int a, int b) { return didx[doff[a]] < didx[doff[b]]; });
		RB_CHECK(s.order.ensure((size_t) P * sizeof(int)));
		RB_CUDA(cudaMemcpyAsync(s.order.p, s.h_order.data(), (size_t) P * sizeof(int), cudaMemcpyHostToDevice, cs));
		s.has_order = true;
	}
	if (priors)
	{
		const size_t nd = pool->dir_off[P], np = pool->psi_off[P];
		RB_CHECK(s.dir_idx.ensure(nd * 4)); RB_CHECK(s.dir_prior.ensure(nd * 8));
		RB_CHECK(s.psi_idx.ensure(np * 4)); RB_CHECK(s.psi_prior.ensure(np * 8));
		RB_CUDA(cudaMemcpyAsync(s.dir_idx.p, pool->dir_idx, nd * 4, cudaMemcpyHostToDevice, cs));
		RB_CUDA(cudaMemcpyAsync(s.dir_prior.p, pool->dir_prior, nd * 8, cudaMemcpyHostToDevice, cs));
		RB_CUDA(cudaMemcpyAsync(s.psi_idx.p, pool->psi_idx, np * 4, cudaMemcpyHostToDevice, cs));
		RB_CUDA(cudaMemcpyAsync(s.psi_prior.p, pool->psi_prior, np * 8, cudaMemcpyHostToDevice, cs));
	}
	s.has_pre_shift = false;
	if (pool->pre_shift)
	{
		RB_CHECK(s.pre_shift.ensure((size_t) P * 2 * sizeof(double)));
		RB_CUDA(cudaMemcpyAsync(s.pre_shift.p, pool->pre_shift, (size_t) P * 2 * sizeof(double), cudaMemcpyHostToDevice, cs));
		s.has_pre_shift = true;
		if (copy_images) RB_CHECK(rbk_pre_shift(ctx, s, cs));     // rb_pool_prepare applies it after its own kernels
	}
	RB_CUDA(cudaEventRecord(s.uploaded, cs));

	// work buffers
	RB_CHECK(s.state.ensure(P * sizeof(RbPartState)));
	RB_CHECK(s.Mweight.ensure((size_t) s.total_coarse * 4));
	RB_CHECK(s.pdf_orient.ensure((size_t) s.total_prior * 4)); RB_CHECK(s.pdf_orient_zero.ensure((size_t) s.total_prior));
	RB_CHECK(s.pdf_offset.ensure((size_t) P * M.prior_classes() * T * 4)); RB_CHECK(s.pdf_offset_zero.ensure((size_t) P * M.prior_classes() * T));
	RB_CHECK(s.counters.ensure(64));
	RB_CHECK(s.shells.ensure((size_t) P * M.nshell * 4));
	RB_CHECK(s.out_pdf_dir.ensure((size_t) K * S.n_dir * 8)); RB_CHECK(s.out_pdf_class.ensure((size_t) 3 * K * 8));   // [K] + [K][2] prior-offset sums
	// fine-pass capacity: all coarse samples significant is the worst case; bound it by a budget
	const long long ov = (long long) S.n_over_rot * S.n_over_trans;
	size_t cap_fs = (size_t) std::min<long long>(s.total_coarse * ov, (long long) env_size("RB_FINE_SAMPLE_CAP", (size_t) 1 << 26));
	size_t cap_fo = (size_t) std::min<long long>(s.total_prior * S.n_over_rot, (long long) env_size("RB_FINE_ORIENT_CAP", (size_t) 1 << 22));
	s.cap_fs = cap_fs; s.cap_fo = cap_fo;
	RB_CHECK(s.fs_w.ensure(cap_fs * 4)); RB_CHECK(s.fs_ihid.ensure(cap_fs * 8));
	// slice cache: the fine pass leaves every projected slice here so the store stage streams it instead of
	// gathering from the reference a second time; orientations beyond the budget fall back to the gather
	{
		const size_t slice_bytes = (size_t) M.Npf * sizeof(float2);
		const size_t budget = env_size("RB_SLICE_CACHE_BYTES", (size_t) 8 << 30);
		// band-major path: one buffer of band-ordered slices per context (grows to the largest request, never shrinks).  The
		// host does not know how many fine orientations the coarse pass will leave, so it launches a fixed number of rounds
		// (RB_BAND_ROUNDS, default 4; rounds beyond the actual count exit at once) and the fine lists are capped accordingly:
		// a pool that needs more reports RB_ERR_CAPACITY like any other fine-pass overflow.
		const size_t band_bytes = (size_t) M.nv_rs_pad * sizeof(float2);
		const long long band_cap = (long long) std::min<size_t>(cap_fo, budget / band_bytes);
		const bool band = !M.do_cc && band_cap > 0 && env_size("RB_BAND", 1) != 0;
		if (band)
		{
			if (band_cap > ctx->band_slice_capacity || ctx->band_slices.bytes < (size_t) band_cap * band_bytes)
			{
				RB_CUDA(cudaStreamSynchronize(ctx->stream));      // the other slot's E-step may still read the old buffer
				RB_CHECK(ctx->band_slices.ensure((size_t) band_cap * band_bytes));
			}
			ctx->band_slice_capacity = band_cap;
			s.slice_capacity = 0;
			const long long max_rounds = (long long) std::max<size_t>(1, env_size("RB_BAND_ROUNDS", 4));
			s.band_rounds = (int) std::min<long long>(((long long) cap_fo + band_cap - 1) / band_cap, max_rounds);
			cap_fo = (size_t) std::min<long long>((long long) cap_fo, s.band_rounds * band_cap);
			s.cap_fo = cap_fo;
		}
		else
		{
			ctx->band_slice_capacity = 0;
			s.band_rounds = 0;
			s.slice_capacity = (long long) std::min<size_t>(cap_fo, budget / slice_bytes);
			RB_CHECK(s.slices.ensure((size_t) s.slice_capacity * slice_bytes));
		}
	}
	RB_CHECK(s.fo.ensure(cap_fo * sizeof(RbFineOrient)));
	RB_CHECK(s.pair_list.ensure((cap_fs / ov + 1) * 4));
	return RB_OK;
}

extern "C" int rb_pool_upload(rb_ctx *ctx, int slot, const rb_particles *pool) { return pool_setup(ctx, slot, pool, true); }

// getFourierTransformsAndCtfs for the whole pool on the device (kernels_prep.cu)
// CTF::initialise (src/ctf.cpp:211-261): the nine constants CTF::getCTF evaluates with (K1 .. K5, the astigmatism matrix, scale)
static void ctf_params(double kV, double Cs, double Q0, double defU, double defV, double defAngle, double Bfac, double phase_shift,
                       double scale, double *q)
{
	const double local_Cs = Cs * 1e7, local_kV = kV * 1e3;
	const double az = defAngle * 3.14159265358979323846 / 180.0;
	const double lam = 12.2643247 / sqrt(local_kV * (1.0 + local_kV * 0.978466e-6));
	q[0] = 3.14159265358979323846 / 2 * 2 * lam;
	q[1] = 3.14159265358979323846 / 2 * local_Cs * lam * lam * lam;
	q[2] = atan(Q0 / sqrt(1 - Q0 * Q0));
	q[3] = -Bfac / 4.0;
	q[4] = phase_shift * 3.14159265358979323846 / 180.0;
	const double ca = cos(az), sa = sin(az), dU = -defU, dV = -defV;
	q[5] = ca * ca * dU + sa * sa * dV;            // A = Q^T D Q, Q = [[ca, sa], [-sa, ca]], D = diag(-defU, -defV)
	q[6] = ca * sa * dU - sa * ca * dV;
	q[7] = sa * sa * dU + ca * ca * dV;
	q[8] = scale;
}

extern "C" int rb_pool_prepare(rb_ctx *ctx, int slot, const rb_raw_particles *raw, float *power_img)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && raw, "rb_pool_prepare: NULL argument");
	if (!ctx->has_model || !ctx->has_sampling) { rb_set_error("rb_pool_prepare: set model and sampling first"); return RB_ERR_STATE; }
	const int P = raw->n_particles, n = raw->image_size;
	RB_ARG(P > 0 && raw->images && raw->old_offset && raw->prior_offset && raw->group_id && raw->optics_group, "rb_pool_prepare: NULL particle array");
	RB_ARG(n == ctx->h_model.ori_size, "rb_pool_prepare: image size %d differs from the model's ori_size %d", n, ctx->h_model.ori_size);
	RB_ARG(n % 2 == 0 && n / 2 + 1 <= 1024, "rb_pool_prepare: image size %d unsupported", n);
	const bool do_ctf = ctx->h_model.do_ctf_correction != 0;
	RB_ARG(!do_ctf || (raw->ctf_defU && raw->ctf_defV && raw->ctf_defAngle && raw->og_kV && raw->og_Cs && raw->og_Q0), "rb_pool_prepare: CTF parameters missing");
	// rounded old offsets (my_old_offset.selfROUND, acc_ml_optimiser_impl.h:216; ROUND of src/macros.h:197)
	std::vector<double> old_r((size_t) 2 * P), xi2((size_t) P, 0.);
	std::vector<int> shift((size_t) 2 * P);
	std::vector<float> norm((size_t) P, 1.f);
	std::vector<double> ctfpar;
	for (int i = 0; i < 2 * P; i++)
	{
		const double v = raw->old_offset[i];
		shift[i] = v > 0 ? (int) (v + 0.5) : (int) (v - 0.5);
		old_r[i] = (double) shift[i];
	}
	if (raw->norm_factor) for (int p = 0; p < P; p++) norm[p] = (float) raw->norm_factor[p];
	if (do_ctf)
	{
		ctfpar.resize((size_t) P * 9);
		for (int p = 0; p < P; p++)
		{
			const int og = raw->optics_group[p];
			RB_ARG(og >= 0 && og < ctx->h_model.nr_optics_groups, "rb_pool_prepare: optics group of particle %d out of range", p);
			ctf_params(raw->og_kV[og], raw->og_Cs[og], raw->og_Q0[og], raw->ctf_defU[p], raw->ctf_defV[p], raw->ctf_defAngle[p],
			           raw->ctf_Bfac ? raw->ctf_Bfac[p] : 0., raw->ctf_phase_shift ? raw->ctf_phase_shift[p] : 0.,
			           raw->ctf_scale ? raw->ctf_scale[p] : 1.0, &ctfpar[(size_t) p * 9]);
		}
	}
	rb_particles pool;
	memset(&pool, 0, sizeof(pool));
	pool.n_particles = P; pool.group_id = raw->group_id; pool.optics_group = raw->optics_group;
	pool.highres_Xi2 = xi2.data(); pool.old_offset = old_r.data(); pool.prior_offset = raw->prior_offset;
	pool.dir_off = raw->dir_off; pool.dir_idx = raw->dir_idx; pool.dir_prior = raw->dir_prior;
	pool.psi_off = raw->psi_off; pool.psi_idx = raw->psi_idx; pool.psi_prior = raw->psi_prior;
	pool.bp_offset = raw->bp_offset;
	pool.mat_left = raw->mat_left; pool.mat_right = raw->mat_right; pool.pre_shift = raw->pre_shift;
	RB_CHECK(pool_setup(ctx, slot, &pool, false));
	PoolSlot &s = ctx->slot[slot];
	// raw images + small tables on the copy stream (overlaps the compute of the other slot), kernels on the compute stream
	DevBuf &bRaw = ctx->prep_raw[slot][0], &bShift = ctx->prep_raw[slot][1], &bNorm = ctx->prep_raw[slot][2], &bCtf = ctx->prep_raw[slot][3];
	DevBuf &bPow = ctx->prep_buf[3];
	const size_t raw_bytes = (size_t) P * n * n * 4;
	RB_CHECK(bRaw.ensure(raw_bytes)); RB_CHECK(bShift.ensure((size_t) P * 8)); RB_CHECK(bNorm.ensure((size_t) P * 4));
	RB_CHECK(bCtf.ensure(std::max<size_t>(ctfpar.size() * 8, 8))); RB_CHECK(bPow.ensure((size_t) P * (n / 2 + 1) * 4));
	cudaStream_t cs = ctx->copy_stream;
	RB_CUDA(cudaMemcpyAsync(bRaw.p, raw->images, raw_bytes, cudaMemcpyHostToDevice, cs));
	RB_CUDA(cudaMemcpyAsync(bShift.p, shift.data(), (size_t) P * 8, cudaMemcpyHostToDevice, cs));
	RB_CUDA(cudaMemcpyAsync(bNorm.p, norm.data(), (size_t) P * 4, cudaMemcpyHostToDevice, cs));
	if (do_ctf) RB_CUDA(cudaMemcpyAsync(bCtf.p, ctfpar.data(), ctfpar.size() * 8, cudaMemcpyHostToDevice, cs));
	// noise-filled soft mask: per-particle seeds and sqrt(sigma2_fudge * sigma2_noise) per optics group (utilities_impl.h:248-249)
	const long long *d_seed = nullptr; const float *d_spec = nullptr;
	std::vector<float> spec;
	if (raw->noise_seed)
	{
		const int nshell = ctx->d_model.nshell, nog = ctx->h_model.nr_optics_groups;
		spec.resize((size_t) nog * nshell);
		const double *s2 = raw->noise_sigma2 ? raw->noise_sigma2 : ctx->h_sigma2_noise.data();
		for (size_t i = 0; i < spec.size(); i++) spec[i] = (float) sqrt(ctx->h_model.sigma2_fudge * s2[i]);
		DevBuf &bSeed = ctx->prep_raw[slot][4], &bSpec = ctx->prep_raw[slot][5];
		RB_CHECK(bSeed.ensure((size_t) P * 8)); RB_CHECK(bSpec.ensure(spec.size() * 4));
		RB_CUDA(cudaMemcpyAsync(bSeed.p, raw->noise_seed, (size_t) P * 8, cudaMemcpyHostToDevice, cs));
		RB_CUDA(cudaMemcpyAsync(bSpec.p, spec.data(), spec.size() * 4, cudaMemcpyHostToDevice, cs));
		d_seed = bSeed.as<long long>(); d_spec = bSpec.as<float>();
	}
	// beam-tilt / MTF factor images of the optics groups
	const float2 *d_fac = nullptr;
	if (raw->og_fourier_factor)
	{
		const int csz = ctx->h_model.current_size;
		const size_t fb = (size_t) ctx->h_model.nr_optics_groups * csz * (csz / 2 + 1) * sizeof(float2);
		DevBuf &bFac = ctx->prep_raw[slot][6];
		RB_CHECK(bFac.ensure(fb));
		RB_CUDA(cudaMemcpyAsync(bFac.p, raw->og_fourier_factor, fb, cudaMemcpyHostToDevice, cs));
		d_fac = bFac.as<float2>();
	}
	// The preparation kernels follow the copies on the COPY stream: they run beside the E-step of the other slot on the compute
	// stream (which is where a RELION adapter spends its time) instead of in front of this slot's E-step.
	RB_CHECK(rbk_prepare_pool(ctx, s, bRaw.as<float>(), bShift.as<int>(), bNorm.as<float>(), do_ctf ? bCtf.as<double>() : nullptr, n,
	                          (float) raw->mask_radius, (float) raw->width_mask_edge, bPow.as<float>(), d_seed, d_spec, d_fac, cs));
	if (power_img) RB_CUDA(cudaMemcpyAsync(power_img, bPow.p, (size_t) P * (n / 2 + 1) * 4, cudaMemcpyDeviceToHost, cs));
	if (s.has_pre_shift) RB_CHECK(rbk_pre_shift(ctx, s, cs));
	RB_CUDA(cudaEventRecord(s.uploaded, cs));            // rb_estep_slot waits for the preparation
	RB_CUDA(cudaStreamSynchronize(cs));                  // shift / norm / ctfpar / spec are host temporaries; power_img is the caller's
	return RB_OK;
}

// read a staged / prepared pool back (tests): any pointer may be NULL
extern "C" int rb_pool_download(rb_ctx *ctx, int slot, float *Fimg, float *Fimg_nomask, float *Fctf, double *highres_Xi2)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && slot >= 0 && slot < RB_NUM_SLOTS && ctx->slot[slot].P > 0, "rb_pool_download: slot %d not staged", slot);
	PoolSlot &s = ctx->slot[slot];
	const size_t img_bytes = (size_t) s.P * ctx->d_model.Npf * sizeof(float2);
	RB_CUDA(cudaStreamWaitEvent(ctx->stream, s.uploaded, 0));
	if (Fimg) RB_CUDA(cudaMemcpyAsync(Fimg, s.Fimg.p, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
	if (Fimg_nomask) RB_CUDA(cudaMemcpyAsync(Fimg_nomask, s.Fnomask.p, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
	if (Fctf && ctx->h_model.do_ctf_correction) RB_CUDA(cudaMemcpyAsync(Fctf, s.Fctf.p, img_bytes / 2, cudaMemcpyDeviceToHost, ctx->stream));
	std::vector<RbPartMeta> m(s.P);
	RB_CUDA(cudaMemcpyAsync(m.data(), s.meta.p, s.P * sizeof(RbPartMeta), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	if (highres_Xi2) for (int p = 0; p < s.P; p++) highres_Xi2[p] = 2.0 * (double) m[p].xi2_half;
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// pool E-step
// ---------------------------------------------------------------------------------------------
static int run_slot(rb_ctx *ctx, PoolSlot &s, unsigned flags)
{
	const RbModelDev &M = ctx->d_model;
	const RbSamplingDev &S = ctx->d_samp;
	for (int k = 0; k < M.nr_classes; k++)
	{
		if (!ctx->has_proj[k]) { rb_set_error("rb_estep: reference %d not set", k); return RB_ERR_STATE; }
		if (!(flags & 1u) && !ctx->has_bp[k]) { rb_set_error("rb_estep: accumulator %d not initialised", k); return RB_ERR_STATE; }
		if (!(flags & 1u) && s.max_bp_off > 0 && !ctx->has_bp[k + s.max_bp_off]) { rb_set_error("rb_estep: accumulator %d (pseudo half-set) not initialised", k + s.max_bp_off); return RB_ERR_STATE; }
		const int half = M.current_size / 2, want = ctx->proj[k].mdlMaxR < half ? ctx->proj[k].mdlMaxR : 0;
		if (want != ctx->fine_dead_maxR)
		{
			rb_set_error("rb_estep: reference %d ends at r_max %d, the image window at %d: rb_model.ref_max_r must be %d (it is %d)", k,
			             ctx->proj[k].mdlMaxR, half, want, ctx->fine_dead_maxR);
			return RB_ERR_STATE;
		}
	}

	RB_CUDA(cudaSetDevice(ctx->device));
	RB_CHECK(ensure_coarse_core(ctx, s.has_priors));
	RB_CUDA(cudaStreamWaitEvent(ctx->stream, s.uploaded, 0));
	if (memcmp(&s.lr, &ctx->coarse_lr, sizeof(RbLR)) != 0)
	{
		// the coarse matrices are shared by the pools; a pool with other MBL / MBR rebuilds them in stream order
		ctx->coarse_lr = s.lr;
		ctx->samp_version++;                                    // invalidates everything derived from the coarse matrices
		RB_CHECK(rbk_make_coarse_eulers(ctx, S.rot, S.tilt, S.psi, S.n_dir, S.n_psi, ctx->coarse_lr, ctx->s_coarse_eulers.as<float>()));
	}
	RB_CHECK(rb_stage_begin(ctx, "total"));
	RB_CUDA(cudaMemsetAsync(s.counters.p, 0, 64, ctx->stream));
	RB_CUDA(cudaMemsetAsync(s.shells.p, 0, (size_t) s.P * M.nshell * 4, ctx->stream));
	RB_CUDA(cudaMemsetAsync(s.out_pdf_dir.p, 0, (size_t) M.nr_classes * S.n_dir * 8, ctx->stream));
	RB_CUDA(cudaMemsetAsync(s.out_pdf_class.p, 0, (size_t) 3 * M.nr_classes * 8, ctx->stream));

	const bool band = rbk_band_applicable(ctx);
	if (band) RB_CHECK(rbk_band_images_async(ctx, s));
	RB_CHECK(rb_stage_begin(ctx, "coarse"));
	RB_CHECK(rbk_prep_priors(ctx, s));
	RB_CHECK(rbk_diff2_coarse_pool(ctx, s));
	RB_CHECK(rb_stage_end(ctx, "coarse"));

	RB_CHECK(rb_stage_begin(ctx, "weights_coarse"));
	RB_CHECK(rbk_weights_coarse_pool(ctx, s));
	RB_CHECK(rb_stage_end(ctx, "weights_coarse"));

	RB_CHECK(rb_stage_begin(ctx, "fine_setup"));
	RB_CHECK(rbk_fine_setup_pool(ctx, s));
	RB_CHECK(rb_stage_end(ctx, "fine_setup"));

	RB_CHECK(rb_stage_begin(ctx, "fine"));
	if (band) RB_CHECK(rbk_band_fine_pool(ctx, s));
	else RB_CHECK(rbk_diff2_fine_pool(ctx, s));
	RB_CHECK(rb_stage_end(ctx, "fine"));

	RB_CHECK(rb_stage_begin(ctx, "weights_fine"));
	RB_CHECK(rbk_weights_fine_pool(ctx, s));
	RB_CHECK(rbk_collect_pool(ctx, s));
	RB_CHECK(rb_stage_end(ctx, "weights_fine"));

	RB_CHECK(rb_stage_begin(ctx, "store"));
	if (!(flags & 1u))
	{
		RB_CHECK(band ? rbk_band_store_pool(ctx, s) : rbk_store_pool(ctx, s));
		if (band) for (int k = 0; k < RB_MAX_CLASSES; k++) if (ctx->has_bp[k]) ctx->bp_blk_dirty[k] = ctx->bp[k].blkvol != nullptr;
	}
	RB_CHECK(rb_stage_end(ctx, "store"));
	RB_CHECK(rb_stage_end(ctx, "total"));
	RB_CUDA(cudaEventRecord(s.done, ctx->stream));
	return RB_OK;
}

static int fetch_slot(rb_ctx *ctx, PoolSlot &s, rb_pool_out *out)
{
	const RbModelDev &M = ctx->d_model;
	const RbSamplingDev &S = ctx->d_samp;
	const int P = s.P, K = M.nr_classes, T = S.n_trans, NOR = S.n_over_rot, NOT = S.n_over_trans;
	std::vector<RbPartState> st(P);
	std::vector<float> shells((size_t) P * M.nshell);
	std::vector<double> pd((size_t) K * S.n_dir), pcl((size_t) 3 * K);
	int counters[16];
	// results travel on their own stream behind the slot's completion event, so the E-step of the other slot (already
	// enqueued on the compute stream) keeps running while these are read
	cudaStream_t fs = ctx->fetch_stream;
	RB_CUDA(cudaStreamWaitEvent(fs, s.done, 0));
	RB_CUDA(cudaMemcpyAsync(st.data(), s.state.p, P * sizeof(RbPartState), cudaMemcpyDeviceToHost, fs));
	RB_CUDA(cudaMemcpyAsync(shells.data(), s.shells.p, shells.size() * 4, cudaMemcpyDeviceToHost, fs));
	RB_CUDA(cudaMemcpyAsync(pd.data(), s.out_pdf_dir.p, pd.size() * 8, cudaMemcpyDeviceToHost, fs));
	RB_CUDA(cudaMemcpyAsync(pcl.data(), s.out_pdf_class.p, (size_t) 3 * K * 8, cudaMemcpyDeviceToHost, fs));
	RB_CUDA(cudaMemcpyAsync(counters, s.counters.p, 64, cudaMemcpyDeviceToHost, fs));
	RB_CUDA(cudaStreamSynchronize(fs));
	if (counters[2])
	{
		rb_set_error("fine-pass workspace too small: %lld orientations / %lld samples needed, capacity %zu / %zu; split the pool or raise "
		             "RB_FINE_ORIENT_CAP / RB_FINE_SAMPLE_CAP (band-major path: RB_SLICE_CACHE_BYTES x RB_BAND_ROUNDS)", ((long long *) counters)[2], ((long long *) counters)[3],
		             s.cap_fo, s.cap_fs);
		return RB_ERR_CAPACITY;
	}
	if (counters[3])
	{
		rb_set_error("store-stage workspace too small: %lld significant fine samples in the pool; split the pool or raise RB_BP_SAMPLE_CAP",
		             ((long long *) counters)[6]);
		return RB_ERR_CAPACITY;
	}
	int status = RB_OK;
	if (out && out->particles)
	{
		for (int p = 0; p < P; p++)
		{
			const RbPartState &q = st[p];
			const RbPartMeta &m = s.h_meta[p];
			rb_particle_out &o = out->particles[p];
			memset(&o, 0, sizeof(o));
			o.nr_significant_coarse = q.nr_sig_coarse;
			o.n_fine_orient = q.n_so * NOR; o.n_fine_samples = q.n_pairs * NOR * NOT; o.n_bp_orient = q.n_bp;
			o.min_diff2_coarse = q.min_diff2; o.sum_weight_coarse = q.csum_weight; o.significant_weight_coarse = q.csig_weight;
			if (q.status != 0)
			{
				if (status == RB_OK)
				{
					status = q.status;
					rb_set_error("particle %d of the pool: %s", p,
					             q.status == RB_ERR_NO_SIGNIFICANT ? "no significant coarse samples (ERRFILTEREDZERO/ERRNOSIGNIFS)" :
					             q.status == RB_ERR_SUMWEIGHT_ZERO ? "sum of fine weights is zero (ERRSUMWEIGHTZERO)" : "failed");
				}
				continue;
			}
			o.best_ihidden_over = q.best_ihid;
			{   // Indices::fineIndexToFineIndices (src/acc/acc_ml_optimiser.h:78-95)
				long long t = q.best_ihid, ov = (long long) NOR * NOT, no = (long long) m.nd * m.np;
				o.best_class = (int) (t / (no * T * ov)); t -= (long long) o.best_class * no * T * ov;
				o.best_idir = (int) (t / ((long long) m.np * T * ov)); t -= (long long) o.best_idir * m.np * T * ov;
				o.best_ipsi = (int) (t / ((long long) T * ov)); t -= (long long) o.best_ipsi * T * ov;
				o.best_itrans = (int) (t / ov); t -= (long long) o.best_itrans * ov;
				o.best_iover_rot = (int) (t / NOT); t -= (long long) o.best_iover_rot * NOT;
				o.best_iover_trans = (int) t;
			}
			o.min_diff2 = (float) q.min_diff2_final; o.max_weight = q.fmax_weight; o.sum_weight = q.fsum_weight;
			o.significant_weight = q.fsig_weight;
			o.pmax = q.fmax_weight / q.fsum_weight;                                          // :2924
			if (o.pmax > 1.f && status == RB_OK) { status = RB_ERR_PMAX; rb_set_error("particle %d: normalised probability > 1", p); }
			o.dLL_nolog = log((double) q.fsum_weight) - q.min_diff2_final;                   // :3574
			double nc = 0.;
			for (int i = 0; i < M.nshell; i++) nc += (double) shells[(size_t) p * M.nshell + i];
			o.wsum_norm_correction = nc;
			o.wsum_XA = q.wsum_XA; o.wsum_AA = q.wsum_AA; o.sumw = q.sumw; o.wsum_sigma2_offset = q.wsum_s2off;
		}
	}
	if (out && out->wsum_sigma2_noise) memcpy(out->wsum_sigma2_noise, shells.data(), shells.size() * 4);
	if (out && out->wsum_pdf_direction) for (size_t i = 0; i < pd.size(); i++) out->wsum_pdf_direction[i] += pd[i];
	if (out && out->wsum_pdf_class) for (int k = 0; k < K; k++) out->wsum_pdf_class[k] += pcl[k];
	if (out && out->wsum_prior_offset_class && ctx->d_model.prior_offset_class)
		for (int k = 0; k < 2 * K; k++) out->wsum_prior_offset_class[k] += pcl[K + k];
	return status;
}

extern "C" int rb_estep_slot_nocopy(rb_ctx *ctx, int slot, unsigned flags)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && slot >= 0 && slot < RB_NUM_SLOTS && ctx->slot[slot].P > 0, "rb_estep_slot_nocopy: slot %d not uploaded", slot);
	return run_slot(ctx, ctx->slot[slot], flags);
}

extern "C" int rb_estep_fetch(rb_ctx *ctx, int slot, rb_pool_out *out)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_ARG(ctx && slot >= 0 && slot < RB_NUM_SLOTS && ctx->slot[slot].P > 0, "rb_estep_fetch: slot %d not uploaded", slot);
	return fetch_slot(ctx, ctx->slot[slot], out);
}

extern "C" int rb_debug_prepared_coarse_image(rb_ctx *ctx, int slot, int particle, float *out)
{
	RB_ARG(ctx && out && slot >= 0 && slot < RB_NUM_SLOTS && ctx->slot[slot].P > 0, "rb_debug_prepared_coarse_image: slot %d not uploaded", slot);
	PoolSlot &s = ctx->slot[slot];
	RB_ARG(particle >= 0 && particle < s.P, "rb_debug_prepared_coarse_image: particle %d out of range", particle);
	RB_CUDA(cudaSetDevice(ctx->device));
	const int nc = ctx->d_model.coarse_size, xs = nc / 2 + 1;
	const size_t bytes = (size_t) nc * xs * sizeof(float4);
	RB_ARG(s.cimg4.bytes >= (size_t) s.P * bytes, "rb_debug_prepared_coarse_image: run the E-step of the slot first");
	RB_CUDA(cudaMemcpyAsync(out, (const char *) s.cimg4.p + (size_t) particle * bytes, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_debug_coarse_eulers(rb_ctx *ctx, float *out, long long capacity)
{
	RB_ARG(ctx && out && ctx->has_sampling, "rb_debug_coarse_eulers: rb_set_sampling first");
	RB_CUDA(cudaSetDevice(ctx->device));
	const long long n = (long long) ctx->d_samp.n_dir * ctx->d_samp.n_psi * 9;
	RB_ARG(capacity >= n, "rb_debug_coarse_eulers: buffer of %lld floats, %lld needed", capacity, n);
	RB_CUDA(cudaMemcpyAsync(out, ctx->s_coarse_eulers.p, (size_t) n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_debug_prep_noise(rb_ctx *ctx, int n_particles, int image_size, float *out)
{
	RB_ARG(ctx && out && n_particles > 0 && image_size > 0, "rb_debug_prep_noise: bad argument");
	RB_CUDA(cudaSetDevice(ctx->device));
	const size_t bytes = (size_t) n_particles * image_size * image_size * 4;
	RB_ARG(ctx->prep_buf[4].bytes >= bytes, "rb_debug_prep_noise: no noise images of that size (rb_pool_prepare with noise_seed first)");
	RB_CUDA(cudaMemcpyAsync(out, ctx->prep_buf[4].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_debug_coarse_weights(rb_ctx *ctx, int slot, int particle, float *out, long long capacity, long long *n_out)
{
	RB_ARG(ctx && slot >= 0 && slot < RB_NUM_SLOTS && ctx->slot[slot].P > 0, "rb_debug_coarse_weights: slot %d not uploaded", slot);
	PoolSlot &s = ctx->slot[slot];
	RB_ARG(particle >= 0 && particle < s.P, "rb_debug_coarse_weights: particle %d out of range", particle);
	const RbPartMeta &m = s.h_meta[particle];
	const long long n = (long long) ctx->d_model.nr_classes * m.nd * m.np * ctx->d_samp.n_trans;
	if (n_out) *n_out = n;
	if (!out) return RB_OK;
	RB_ARG(capacity >= n, "rb_debug_coarse_weights: buffer of %lld floats, %lld needed", capacity, n);
	RB_CUDA(cudaSetDevice(ctx->device));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	RB_CUDA(cudaMemcpy(out, s.Mweight.as<float>() + m.coarse_off, (size_t) n * 4, cudaMemcpyDeviceToHost));
	return RB_OK;
}

extern "C" int rb_estep_slot(rb_ctx *ctx, int slot, rb_pool_out *out, unsigned flags)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_CHECK(rb_estep_slot_nocopy(ctx, slot, flags));
	return fetch_slot(ctx, ctx->slot[slot], out);
}

extern "C" int rb_estep_pool(rb_ctx *ctx, const rb_particles *pool, rb_pool_out *out, unsigned flags)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_CHECK(rb_pool_upload(ctx, 0, pool));
	return rb_estep_slot(ctx, 0, out, flags);
}

// ---------------------------------------------------------------------------------------------
// stage-level entry points (host pointers in/out)
// ---------------------------------------------------------------------------------------------
struct StageBufs {
	rb_ctx *ctx; int next = 0; std::vector<DevBuf> owned;
	explicit StageBufs(rb_ctx *c) : ctx(c) {}
	~StageBufs() { for (auto &b : owned) b.release(); }
	template <typename T> int up(const T *src, size_t n, T **dst)
	{
		owned.emplace_back();
		DevBuf &b = owned.back();
		RB_CHECK(b.ensure(std::max<size_t>(n, 1) * sizeof(T)));
		if (src && n) RB_CUDA(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
		else RB_CUDA(cudaMemsetAsync(b.p, 0, std::max<size_t>(n, 1) * sizeof(T), ctx->stream));
		*dst = b.as<T>();
		return RB_OK;
	}
};

static int check_proj(rb_ctx *ctx, int k, int img_size)
{
	RB_ARG(ctx, "ctx is NULL");
	RB_ARG(k >= 0 && k < RB_MAX_CLASSES && ctx->has_proj[k], "reference %d not set", k);
	RB_ARG(img_size > 0 && img_size % 2 == 0 && img_size <= 1000, "image size %d unsupported", img_size);
	RB_CUDA(cudaSetDevice(ctx->device));
	return RB_OK;
}

extern "C" int rb_project(rb_ctx *ctx, int k, int n, const float *eulers, int count, float *out_complex)
{
	if (ctx) cudaSetDevice(ctx->device);   // the caller's thread may have another device current
	RB_CHECK(check_proj(ctx, k, n));
	StageBufs sb(ctx);
	float *d_e; float2 *d_o;
	const size_t np = (size_t) n * (n / 2 + 1);
	RB_CHECK(sb.up(eulers, (size_t) count * 9, &d_e));
	RB_CHECK(sb.up((const float2 *) nullptr, np * count, &d_o));
	RB_CHECK(rbk_project(ctx, ctx->proj[k], n, d_e, count, d_o));
	RB_CUDA(cudaMemcpyAsync(out_complex, d_o, np * count * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_gemm_tf32x3(rb_ctx *ctx, const float *A, const float *B, int M, int N, int K, float *Cout)
{
	RB_ARG(ctx && A && B && Cout, "rb_gemm_tf32x3: NULL argument");
	RB_ARG(M > 0 && N > 0 && K > 0, "rb_gemm_tf32x3: empty problem %d x %d x %d", M, N, K);
	RB_CUDA(cudaSetDevice(ctx->device));
	StageBufs sb(ctx);
	float *d_a, *d_b, *d_c;
	RB_CHECK(sb.up(A, (size_t) M * K, &d_a)); RB_CHECK(sb.up(B, (size_t) N * K, &d_b));
	RB_CHECK(sb.up((const float *) nullptr, (size_t) M * N, &d_c));
	RB_CHECK(rbk_gemm_tf32x3_stage(ctx, d_a, d_b, M, N, K, d_c));
	RB_CUDA(cudaMemcpyAsync(Cout, d_c, (size_t) M * N * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

static int diff2_coarse_entry(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                              const float *tx, const float *ty, int T,
                              const float *re, const float *im, const float *corr, float *diff2s, int cc)
{
	RB_CHECK(check_proj(ctx, k, n));
	StageBufs sb(ctx);
	const size_t np = (size_t) n * (n / 2 + 1);
	float *d_e, *d_tx, *d_ty, *d_re, *d_im, *d_c, *d_o;
	RB_CHECK(sb.up(eulers, (size_t) O * 9, &d_e)); RB_CHECK(sb.up(tx, T, &d_tx)); RB_CHECK(sb.up(ty, T, &d_ty));
	RB_CHECK(sb.up(re, np, &d_re)); RB_CHECK(sb.up(im, np, &d_im)); RB_CHECK(sb.up(corr, np, &d_c));
	RB_CHECK(sb.up(diff2s, (size_t) O * T, &d_o));
	RB_CHECK(rbk_diff2_coarse_stage(ctx, proj_full(ctx, k), n, d_e, O, d_tx, d_ty, T, d_re, d_im, d_c, d_o, cc));
	RB_CUDA(cudaMemcpyAsync(diff2s, d_o, (size_t) O * T * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_diff2_coarse(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                               const float *tx, const float *ty, int T,
                               const float *re, const float *im, const float *corr, float *diff2s)
{
	return diff2_coarse_entry(ctx, k, n, eulers, O, tx, ty, T, re, im, corr, diff2s, 0);
}

extern "C" int rb_diff2_cc_coarse(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                                  const float *tx, const float *ty, int T,
                                  const float *re, const float *im, const float *corr, float *diff2s)
{
	return diff2_coarse_entry(ctx, k, n, eulers, O, tx, ty, T, re, im, corr, diff2s, 1);
}

static int diff2_fine_entry(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                            const float *tx, const float *ty, int T,
                            const float *re, const float *im, const float *corr, float sum_init,
                            const uint64_t *rot_idx, const uint64_t *trans_idx,
                            const uint64_t *job_idx, const uint64_t *job_num, int n_jobs,
                            float *diff2s, int n_weights, int cc)
{
	RB_CHECK(check_proj(ctx, k, n));
	for (int j = 0; j < n_jobs; j++) RB_ARG(job_num[j] <= 16, "rb_diff2_fine: job %d has %llu translations (max 16)", j, (unsigned long long) job_num[j]);
	StageBufs sb(ctx);
	const size_t np = (size_t) n * (n / 2 + 1);
	float *d_e, *d_tx, *d_ty, *d_re, *d_im, *d_c, *d_o;
	unsigned long long *d_ri, *d_ti, *d_ji, *d_jn;
	RB_CHECK(sb.up(eulers, (size_t) O * 9, &d_e)); RB_CHECK(sb.up(tx, T, &d_tx)); RB_CHECK(sb.up(ty, T, &d_ty));
	RB_CHECK(sb.up(re, np, &d_re)); RB_CHECK(sb.up(im, np, &d_im)); RB_CHECK(sb.up(corr, np, &d_c));
	RB_CHECK(sb.up(diff2s, (size_t) n_weights, &d_o));
	RB_CHECK(sb.up((const unsigned long long *) rot_idx, (size_t) n_weights, &d_ri));
	RB_CHECK(sb.up((const unsigned long long *) trans_idx, (size_t) n_weights, &d_ti));
	RB_CHECK(sb.up((const unsigned long long *) job_idx, (size_t) n_jobs, &d_ji));
	RB_CHECK(sb.up((const unsigned long long *) job_num, (size_t) n_jobs, &d_jn));
	RB_CHECK(rbk_diff2_fine_stage(ctx, ctx->proj[k], n, d_e, d_tx, d_ty, d_re, d_im, d_c, sum_init, d_ri, d_ti, d_ji, d_jn, n_jobs, d_o, cc));
	RB_CUDA(cudaMemcpyAsync(diff2s, d_o, (size_t) n_weights * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_diff2_fine(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                             const float *tx, const float *ty, int T,
                             const float *re, const float *im, const float *corr, float sum_init,
                             const uint64_t *rot_idx, const uint64_t *trans_idx,
                             const uint64_t *job_idx, const uint64_t *job_num, int n_jobs,
                             float *diff2s, int n_weights)
{
	return diff2_fine_entry(ctx, k, n, eulers, O, tx, ty, T, re, im, corr, sum_init, rot_idx, trans_idx, job_idx, job_num, n_jobs,
	                        diff2s, n_weights, 0);
}

extern "C" int rb_diff2_cc_fine(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                                const float *tx, const float *ty, int T,
                                const float *re, const float *im, const float *corr,
                                const uint64_t *rot_idx, const uint64_t *trans_idx,
                                const uint64_t *job_idx, const uint64_t *job_num, int n_jobs,
                                float *diff2s, int n_weights)
{
	return diff2_fine_entry(ctx, k, n, eulers, O, tx, ty, T, re, im, corr, 0.f, rot_idx, trans_idx, job_idx, job_num, n_jobs,
	                        diff2s, n_weights, 1);
}

extern "C" int rb_convert_weights(rb_ctx *ctx, float *weights, int64_t n_orient, int n_trans,
                                  const float *pdf_o, const unsigned char *pdf_oz,
                                  const float *pdf_t, const unsigned char *pdf_tz,
                                  double adaptive_fraction, int maxsig, int filter_zero,
                                  unsigned char *significant, rb_weights_out *out)
{
	RB_ARG(ctx && weights && pdf_o && pdf_oz && pdf_t && pdf_tz && out, "rb_convert_weights: NULL argument");
	RB_ARG(n_orient > 0 && n_trans > 0, "rb_convert_weights: empty input");
	RB_CUDA(cudaSetDevice(ctx->device));
	StageBufs sb(ctx);
	const size_t n = (size_t) n_orient * n_trans;
	float *d_w, *d_po, *d_pt; unsigned char *d_oz, *d_tz, *d_sig; rb_weights_out *d_out;
	RB_CHECK(sb.up(weights, n, &d_w)); RB_CHECK(sb.up(pdf_o, (size_t) n_orient, &d_po)); RB_CHECK(sb.up(pdf_t, (size_t) n_trans, &d_pt));
	RB_CHECK(sb.up(pdf_oz, (size_t) n_orient, &d_oz)); RB_CHECK(sb.up(pdf_tz, (size_t) n_trans, &d_tz));
	RB_CHECK(sb.up((const unsigned char *) nullptr, n, &d_sig)); RB_CHECK(sb.up((const rb_weights_out *) nullptr, 1, &d_out));
	RB_CHECK(rbk_convert_weights_stage(ctx, d_w, n_orient, n_trans, d_po, d_oz, d_pt, d_tz, adaptive_fraction, maxsig, filter_zero, d_sig, d_out));
	RB_CUDA(cudaMemcpyAsync(weights, d_w, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (significant) RB_CUDA(cudaMemcpyAsync(significant, d_sig, n, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(out, d_out, sizeof(rb_weights_out), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_wavg(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                       const float *tx, const float *ty, int T,
                       const float *re, const float *im, const float *weights, const float *ctfs,
                       float weight_norm, float sig_w, float *parts, float *AA, float *XA)
{
	RB_CHECK(check_proj(ctx, k, n));
	StageBufs sb(ctx);
	const size_t np = (size_t) n * (n / 2 + 1);
	float *d_e, *d_tx, *d_ty, *d_re, *d_im, *d_w, *d_c, *d_p, *d_a, *d_x;
	RB_CHECK(sb.up(eulers, (size_t) O * 9, &d_e)); RB_CHECK(sb.up(tx, T, &d_tx)); RB_CHECK(sb.up(ty, T, &d_ty));
	RB_CHECK(sb.up(re, np, &d_re)); RB_CHECK(sb.up(im, np, &d_im)); RB_CHECK(sb.up(weights, (size_t) O * T, &d_w));
	RB_CHECK(sb.up(ctfs, np, &d_c)); RB_CHECK(sb.up(parts, np, &d_p)); RB_CHECK(sb.up(AA, np, &d_a)); RB_CHECK(sb.up(XA, np, &d_x));
	RB_CHECK(rbk_wavg_stage(ctx, ctx->proj[k], n, d_e, O, d_tx, d_ty, T, d_re, d_im, d_w, d_c, weight_norm, sig_w, d_p, d_a, d_x));
	RB_CUDA(cudaMemcpyAsync(parts, d_p, np * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(AA, d_a, np * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(XA, d_x, np * 4, cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_backproject(rb_ctx *ctx, int k, int n, const float *eulers, int O,
                              const float *tx, const float *ty, int T,
                              const float *re, const float *im,
                              const float *weights, const float *Minvsigma2s, const float *ctfs,
                              float weight_norm, float sig_w)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_backproject: accumulator %d not initialised", k);
	RB_ARG(n > 0 && n % 2 == 0 && n <= 1000, "image size %d unsupported", n);
	RB_CUDA(cudaSetDevice(ctx->device));
	StageBufs sb(ctx);
	const size_t np = (size_t) n * (n / 2 + 1);
	float *d_e, *d_tx, *d_ty, *d_re, *d_im, *d_w, *d_m, *d_c;
	RB_CHECK(sb.up(eulers, (size_t) O * 9, &d_e)); RB_CHECK(sb.up(tx, T, &d_tx)); RB_CHECK(sb.up(ty, T, &d_ty));
	RB_CHECK(sb.up(re, np, &d_re)); RB_CHECK(sb.up(im, np, &d_im)); RB_CHECK(sb.up(weights, (size_t) O * T, &d_w));
	RB_CHECK(sb.up(Minvsigma2s, np, &d_m)); RB_CHECK(sb.up(ctfs, np, &d_c));
	const int circle = ctx->has_model ? ctx->h_model.bp_circle_bound : 1;
	const int premult = ctx->has_model ? ctx->h_model.ctf_premultiplied : 0;
	RB_CHECK(rbk_backproject_stage(ctx, ctx->bp[k], n, d_e, O, d_tx, d_ty, T, d_re, d_im, d_w, d_m, d_c, weight_norm, sig_w, circle, premult));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

// relion_reconstruct-style posed back-projection: one orientation and unit weight per image, F2D already
// CTF-multiplied, Fctf = ctf^2 (src/reconstructor.cpp:632-737).  Expressed through the same scatter kernel:
// img = F2D, ctf = 1, Minvsigma2 = Fctf  =>  F = F2D, Fweight = Fctf.  r_max and the skipped x=0,y<0 half
// column follow BackProjector::backproject2Dto3D (src/backprojector.cpp:90-91, 150-160).
// Host images in chunks through two device buffers: the H2D copy of chunk i+1 (copy stream) overlaps the scatter of chunk i.
extern "C" int rb_backproject_posed(rb_ctx *ctx, int k, int n, int count,
                                    const float *F2D_complex, const float *Fctf, const float *eulers)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_backproject_posed: accumulator %d not initialised", k);
	RB_ARG(n > 0 && n % 2 == 0 && n <= 1000 && count > 0, "rb_backproject_posed: bad sizes");
	RB_ARG(F2D_complex && Fctf && eulers, "rb_backproject_posed: NULL argument");
	RB_CUDA(cudaSetDevice(ctx->device));
	const size_t np = (size_t) n * (n / 2 + 1);
	const int chunk = (int) std::max<size_t>(1, std::min<size_t>((size_t) count, ((size_t) 512 << 20) / (np * 12)));
	for (int b = 0; b < 2; b++)
	{
		RB_CHECK(ctx->posed_buf[b][0].ensure((size_t) chunk * np * 8)); RB_CHECK(ctx->posed_buf[b][1].ensure((size_t) chunk * np * 4));
		RB_CHECK(ctx->posed_buf[b][2].ensure((size_t) chunk * 36));
		if (!ctx->posed_ev[b]) RB_CUDA(cudaEventCreateWithFlags(&ctx->posed_ev[b], cudaEventDisableTiming));
	}
	ctx->posed_count = 0;   // the staging buffers are being reused
	cudaEvent_t uploaded;
	RB_CUDA(cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming));
	int ib = 0;
	for (int i0 = 0; i0 < count; i0 += chunk, ib ^= 1)
	{
		const int c = std::min(chunk, count - i0);
		// the kernel that last read this buffer must have finished before it is overwritten
		RB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->posed_ev[ib], 0));
		RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[ib][0].p, F2D_complex + (size_t) i0 * np * 2, (size_t) c * np * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
		RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[ib][1].p, Fctf + (size_t) i0 * np, (size_t) c * np * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
		RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[ib][2].p, eulers + (size_t) i0 * 9, (size_t) c * 36, cudaMemcpyHostToDevice, ctx->copy_stream));
		RB_CUDA(cudaEventRecord(uploaded, ctx->copy_stream));
		RB_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded, 0));
		RB_CHECK(rbk_backproject_posed(ctx, ctx->bp[k], n, c, ctx->posed_buf[ib][0].as<float2>(), ctx->posed_buf[ib][1].as<float>(), ctx->posed_buf[ib][2].as<float>()));
		ctx->bp_blk_dirty[k] = ctx->bp[k].blkvol != nullptr;
		RB_CUDA(cudaEventRecord(ctx->posed_ev[ib], ctx->stream));
	}
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	RB_CUDA(cudaEventDestroy(uploaded));
	return RB_OK;
}

// relion_reconstruct from RAW images: FFT, centring, origin shift, CTF and DC removal on the device (rbk_backproject_posed_raw)
extern "C" int rb_backproject_posed_raw(rb_ctx *ctx, int k, const rb_posed_raw *r)
{
	RB_ARG(ctx && r && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_backproject_posed_raw: accumulator %d not initialised", k);
	const int n = r->image_size, count = r->n_images;
	RB_ARG(n > 0 && n % 2 == 0 && n <= 1000 && count > 0, "rb_backproject_posed_raw: bad sizes");
	RB_ARG(r->images && r->eulers, "rb_backproject_posed_raw: NULL argument");
	const bool do_ctf = r->ctf_defU != nullptr;
	RB_ARG(!do_ctf || (r->ctf_defV && r->ctf_defAngle && r->og_kV && r->og_Cs && r->og_Q0 && r->pixel_size > 0.), "rb_backproject_posed_raw: CTF parameters missing");
	RB_CUDA(cudaSetDevice(ctx->device));
	const size_t npx = (size_t) n * n;
	// 128 MB chunks: the copy of chunk i + 1 runs under the transform + scatter of chunk i
	const int chunk = (int) std::max<size_t>(1, std::min<size_t>((size_t) count, ((size_t) 128 << 20) / (npx * 4)));
	std::vector<double> par((size_t) (do_ctf ? count : 0) * 9);
	for (int i = 0; i < count && do_ctf; i++)
	{
		const int og = r->optics_group ? r->optics_group[i] : 0;
		ctf_params(r->og_kV[og], r->og_Cs[og], r->og_Q0[og], r->ctf_defU[i], r->ctf_defV[i], r->ctf_defAngle[i], r->ctf_Bfac ? r->ctf_Bfac[i] : 0.,
		           r->ctf_phase_shift ? r->ctf_phase_shift[i] : 0., r->ctf_scale ? r->ctf_scale[i] : 1.0, &par[(size_t) i * 9]);
	}
	// staging: [0] images, [1] ctf parameters + shifts (doubles), [2] matrices; two sets for copy / compute overlap
	for (int b = 0; b < 2; b++)
	{
		RB_CHECK(ctx->posed_buf[b][0].ensure((size_t) chunk * npx * 4)); RB_CHECK(ctx->posed_buf[b][1].ensure((size_t) chunk * 11 * 8));
		RB_CHECK(ctx->posed_buf[b][2].ensure((size_t) chunk * 36));
		if (!ctx->posed_ev[b]) RB_CUDA(cudaEventCreateWithFlags(&ctx->posed_ev[b], cudaEventDisableTiming));
	}
	ctx->posed_count = 0;
	cudaEvent_t uploaded;
	RB_CUDA(cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming));
	int ib = 0;
	for (int i0 = 0; i0 < count; i0 += chunk, ib ^= 1)
	{
		const int c = std::min(chunk, count - i0);
		RB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->posed_ev[ib], 0));
		double *d_par = ctx->posed_buf[ib][1].as<double>(), *d_shift = d_par + (size_t) chunk * 9;
		RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[ib][0].p, r->images + (size_t) i0 * npx, (size_t) c * npx * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
		if (do_ctf) RB_CUDA(cudaMemcpyAsync(d_par, par.data() + (size_t) i0 * 9, (size_t) c * 72, cudaMemcpyHostToDevice, ctx->copy_stream));
		if (r->shift) RB_CUDA(cudaMemcpyAsync(d_shift, r->shift + (size_t) i0 * 2, (size_t) c * 16, cudaMemcpyHostToDevice, ctx->copy_stream));
		RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[ib][2].p, r->eulers + (size_t) i0 * 9, (size_t) c * 36, cudaMemcpyHostToDevice, ctx->copy_stream));
		RB_CUDA(cudaEventRecord(uploaded, ctx->copy_stream));
		RB_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded, 0));
		RB_CHECK(rbk_backproject_posed_raw(ctx, ctx->bp[k], n, c, ctx->posed_buf[ib][0].as<float>(), r->shift ? d_shift : nullptr,
		                                   do_ctf ? d_par : nullptr, (double) n * r->pixel_size, r->ctf_premultiplied, ctx->posed_buf[ib][2].as<float>()));
		ctx->bp_blk_dirty[k] = ctx->bp[k].blkvol != nullptr;
		RB_CUDA(cudaEventRecord(ctx->posed_ev[ib], ctx->stream));
	}
	RB_CUDA(cudaStreamSynchronize(ctx->copy_stream));      // par is a host temporary
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	RB_CUDA(cudaEventDestroy(uploaded));
	return RB_OK;
}

// Device-resident variant for roofline measurement: stage a batch once, then scatter it any number of times.
extern "C" int rb_bp_posed_stage(rb_ctx *ctx, int n, int count, const float *F2D_complex, const float *Fctf, const float *eulers)
{
	RB_ARG(ctx && F2D_complex && Fctf && eulers, "rb_bp_posed_stage: NULL argument");
	RB_ARG(n > 0 && n % 2 == 0 && n <= 1000 && count > 0, "rb_bp_posed_stage: bad sizes");
	RB_CUDA(cudaSetDevice(ctx->device));
	const size_t np = (size_t) n * (n / 2 + 1);
	RB_CHECK(ctx->posed_buf[0][0].ensure((size_t) count * np * 8)); RB_CHECK(ctx->posed_buf[0][1].ensure((size_t) count * np * 4));
	RB_CHECK(ctx->posed_buf[0][2].ensure((size_t) count * 36));
	RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[0][0].p, F2D_complex, (size_t) count * np * 8, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[0][1].p, Fctf, (size_t) count * np * 4, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(ctx->posed_buf[0][2].p, eulers, (size_t) count * 36, cudaMemcpyHostToDevice, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->posed_n = n; ctx->posed_count = count;
	return RB_OK;
}

extern "C" int rb_bp_posed_run(rb_ctx *ctx, int k)
{
	RB_ARG(ctx && k >= 0 && k < RB_MAX_CLASSES && ctx->has_bp[k], "rb_bp_posed_run: accumulator %d not initialised", k);
	if (ctx->posed_count < 1) { rb_set_error("rb_bp_posed_run: nothing staged (rb_bp_posed_stage)"); return RB_ERR_STATE; }
	RB_CUDA(cudaSetDevice(ctx->device));
	ctx->bp_blk_dirty[k] = ctx->bp[k].blkvol != nullptr;
	return rbk_backproject_posed(ctx, ctx->bp[k], ctx->posed_n, ctx->posed_count, ctx->posed_buf[0][0].as<float2>(),
	                             ctx->posed_buf[0][1].as<float>(), ctx->posed_buf[0][2].as<float>());
}
