// relion_b200 — per-iteration reduction over the GPUs of one box through NCCL, behind the C-ABI.
//
// Replaces MlOptimiserMpi::combineAllWeightedSums (/root/reference/src/ml_optimiser_mpi.cpp:2028-2185), which packs
// MlWsumModel (src/ml_model.cpp:1881-1957), sends it around a ring of MPI ranks (or through files) and unpacks it:
//   rb_bp_allreduce    every class' back-projection accumulator, summed in place on the device, on the context's own
//                      stream (ordered after the E-step kernels, before any later rb_reconstruct / rb_bp_get).  The
//                      accumulator is float4 (re, im, weight, pad): the pad lane does not travel - three planar floats per
//                      voxel are packed into a staging buffer, reduced with ncclAllReduce(ncclFloat, ncclSum) over
//                      NVLink / NVSwitch, and scattered back.
//   rb_wsum_allreduce  the remaining weighted sums as ONE fp64 vector in MlWsumModel::pack order (the C++ adapter's
//                      relion_b200::WsumPack lays it out), staged through device memory, ncclAllReduce(ncclDouble, ncclSum).
// NCCL is loaded at run time (dlopen "libnccl.so.2", or RB_NCCL_LIB): the library has no link-time dependency on it, and a
// process that already holds a copy (torch's bundled one) shares it.  No MPI anywhere.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>

#include <nccl.h>   // types and prototypes only: the functions are resolved with dlsym below

namespace {

struct NcclApi {
	void *handle = nullptr;
	decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
	decltype(&ncclCommInitRank) CommInitRank = nullptr;
	decltype(&ncclCommInitAll) CommInitAll = nullptr;
	decltype(&ncclCommDestroy) CommDestroy = nullptr;
	decltype(&ncclAllReduce) AllReduce = nullptr;
	decltype(&ncclGroupStart) GroupStart = nullptr;
	decltype(&ncclGroupEnd) GroupEnd = nullptr;
	decltype(&ncclGetErrorString) GetErrorString = nullptr;
	decltype(&ncclCommCount) CommCount = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load()
{
	std::lock_guard<std::mutex> lk(g_nccl_mu);
	if (g_nccl.handle) return RB_OK;
	const char *names[] = {getenv("RB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	void *h = nullptr;
	for (const char *n : names) { if (n && *n) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; } }
	if (!h) { rb_set_error("NCCL not found (dlopen libnccl.so.2: %s); set RB_NCCL_LIB", dlerror()); return RB_ERR_STATE; }
#define RB_SYM(field, name) do { *(void **) (&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { rb_set_error("NCCL symbol %s missing", name); dlclose(h); return RB_ERR_STATE; } } while (0)
	RB_SYM(GetUniqueId, "ncclGetUniqueId"); RB_SYM(CommInitRank, "ncclCommInitRank"); RB_SYM(CommInitAll, "ncclCommInitAll");
	RB_SYM(CommDestroy, "ncclCommDestroy"); RB_SYM(AllReduce, "ncclAllReduce"); RB_SYM(GroupStart, "ncclGroupStart");
	RB_SYM(GroupEnd, "ncclGroupEnd"); RB_SYM(GetErrorString, "ncclGetErrorString"); RB_SYM(CommCount, "ncclCommCount");
#undef RB_SYM
	g_nccl.handle = h;
	return RB_OK;
}

#define RB_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { rb_set_error("NCCL error %d at %s:%d: %s", (int) r__, __FILE__, __LINE__, g_nccl.GetErrorString(r__)); return RB_ERR_CUDA; } } while (0)

__global__ void k_bp_pack3(const float4 *vol, float *planar, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const float4 v = vol[i];
		planar[i] = v.x; planar[n + i] = v.y; planar[2 * n + i] = v.z;
	}
}
__global__ void k_bp_unpack3(float4 *vol, const float *planar, size_t n)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
		vol[i] = make_float4(planar[i], planar[n + i], planar[2 * n + i], 0.f);
}

} // namespace

struct rb_comm { ncclComm_t comm; int nranks, rank, device; };

extern "C" int rb_comm_unique_id(void *id128)
{
	RB_ARG(id128, "rb_comm_unique_id: NULL argument");
	RB_CHECK(nccl_load());
	ncclUniqueId id;
	RB_NCCL(g_nccl.GetUniqueId(&id));
	memcpy(id128, &id, sizeof(id));
	return RB_OK;
}

extern "C" int rb_comm_create(rb_ctx *ctx, int nranks, int rank, const void *id128, rb_comm **out)
{
	RB_ARG(ctx && id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, "rb_comm_create: bad argument");
	RB_CHECK(nccl_load());
	RB_CUDA(cudaSetDevice(ctx->device));
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	rb_comm *c = new rb_comm();
	c->nranks = nranks; c->rank = rank; c->device = ctx->device;
	ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
	if (r != ncclSuccess) { rb_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return RB_ERR_CUDA; }
	*out = c;
	return RB_OK;
}

// all ranks inside one process (RELION drives several GPUs from the threads of one process): one communicator per context
extern "C" int rb_comm_create_all(rb_ctx *const *ctxs, int n, rb_comm **out)
{
	RB_ARG(ctxs && out && n >= 1 && n <= 64, "rb_comm_create_all: bad argument");
	RB_CHECK(nccl_load());
	int devs[64]; ncclComm_t comms[64];
	for (int i = 0; i < n; i++) { RB_ARG(ctxs[i], "rb_comm_create_all: NULL context"); devs[i] = ctxs[i]->device; }
	RB_NCCL(g_nccl.CommInitAll(comms, n, devs));
	for (int i = 0; i < n; i++) { out[i] = new rb_comm(); out[i]->comm = comms[i]; out[i]->nranks = n; out[i]->rank = i; out[i]->device = devs[i]; }
	return RB_OK;
}

extern "C" void rb_comm_destroy(rb_comm *c)
{
	if (!c) return;
	if (g_nccl.handle && c->comm) { cudaSetDevice(c->device); g_nccl.CommDestroy(c->comm); }
	delete c;
}

extern "C" int rb_comm_size(const rb_comm *c) { return c ? c->nranks : 0; }
extern "C" int rb_comm_rank(const rb_comm *c) { return c ? c->rank : -1; }
extern "C" int rb_comm_group_start(void) { RB_CHECK(nccl_load()); RB_NCCL(g_nccl.GroupStart()); return RB_OK; }
extern "C" int rb_comm_group_end(void) { RB_CHECK(nccl_load()); RB_NCCL(g_nccl.GroupEnd()); return RB_OK; }

// sum of `nr_classes` accumulators over the ranks of `nccl_comm` (an ncclComm_t: rb_comm_handle(), or the caller's own), in place
extern "C" int rb_bp_allreduce_nccl(rb_ctx *ctx, void *nccl_comm, int nr_classes, int wait)
{
	RB_ARG(ctx && nccl_comm && nr_classes >= 1 && nr_classes <= RB_MAX_CLASSES, "rb_bp_allreduce: bad argument");
	RB_CHECK(nccl_load());
	RB_CUDA(cudaSetDevice(ctx->device));
	ncclComm_t comm = (ncclComm_t) nccl_comm;
	int nranks = 1;
	RB_NCCL(g_nccl.CommCount(comm, &nranks));
	if (nranks > 1)
	{
		for (int k = 0; k < nr_classes; k++)
		{
			RB_ARG(ctx->has_bp[k], "rb_bp_allreduce: accumulator %d not initialised", k);
			RB_CHECK(rb_bp_fold(ctx, k));
			const RbBackprojector &b = ctx->bp[k];
			const size_t n = (size_t) b.mdlX * b.mdlY * b.mdlZ;
			RB_CHECK(ctx->comm_buf.ensure(3 * n * sizeof(float)));
			float *planar = ctx->comm_buf.as<float>();
			k_bp_pack3<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(b.vol, planar, n);
			RB_LAUNCH_CHECK(ctx);
			RB_NCCL(g_nccl.AllReduce(planar, planar, 3 * n, ncclFloat, ncclSum, comm, ctx->stream));
			k_bp_unpack3<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(b.vol, planar, n);
			RB_LAUNCH_CHECK(ctx);
		}
	}
	if (wait) RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}

extern "C" int rb_bp_allreduce(rb_ctx *ctx, rb_comm *comm)
{
	RB_ARG(ctx && comm, "rb_bp_allreduce: NULL argument");
	RB_ARG(ctx->has_model, "rb_bp_allreduce: rb_set_model first (number of classes)");
	return rb_bp_allreduce_nccl(ctx, comm->comm, ctx->h_model.nr_classes, 1);
}

extern "C" void *rb_comm_handle(rb_comm *c) { return c ? (void *) c->comm : nullptr; }

// fp64 vector of weighted sums (MlWsumModel::pack order without the volumes), summed over the ranks; host in / out
extern "C" int rb_wsum_allreduce(rb_ctx *ctx, rb_comm *comm, double *wsums, size_t n)
{
	RB_ARG(ctx && comm && (wsums || n == 0), "rb_wsum_allreduce: NULL argument");
	if (n == 0 || comm->nranks == 1) return RB_OK;
	RB_CHECK(nccl_load());
	RB_CUDA(cudaSetDevice(ctx->device));
	RB_CHECK(ctx->comm_buf2.ensure(n * sizeof(double)));
	RB_CUDA(cudaMemcpyAsync(ctx->comm_buf2.p, wsums, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	RB_NCCL(g_nccl.AllReduce(ctx->comm_buf2.p, ctx->comm_buf2.p, n, ncclDouble, ncclSum, comm->comm, ctx->stream));
	RB_CUDA(cudaMemcpyAsync(wsums, ctx->comm_buf2.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	RB_CUDA(cudaStreamSynchronize(ctx->stream));
	return RB_OK;
}
