// MRC image stacks and the prefetching particle feed: the step before image preparation.
//
// Reader semantics follow the reference (src/rwMRC.h:82-283): 1024-byte header, byte order detected from the magnitude of
// `mode` / `nx` (SWAPTRIG, src/image.h:386) with every 4-byte header word before the labels swapped, data at
// 1024 + nsymbt, modes 0 (SIGNED 8-bit since RELION 3.1), 1 (int16), 2 (float32), 6 (uint16), 12 (float16); complex modes
// 3 / 4 and SerialEM's 4-bit mode 101 are refused.  An .mrcs file is a stack of nz images of nx * ny pixels, x fastest.
// Writer: mode 2, mx/my/mz = nx/ny/nz, cell = n * pixel size, 90-degree angles, axes 1/2/3, statistics, "MAP " and the
// little-endian machine stamp (rwMRC.h:286-530).
//
// The feed replaces the image half of MlOptimiser::getMetaAndImageDataSubset (src/ml_optimiser.cpp:10285-10406): for the
// NEXT pools it gathers the particles' images from their stacks ("only open new stacks": one descriptor per stack, kept
// open, positional reads) into page-locked buffers that rb_pool_prepare uploads with one asynchronous copy, while the GPU
// works on the current pool.  Reader threads split a pool's images between them.
#include "relion_b200.h"
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

void rb_set_error(const char *fmt, ...);

namespace {

const int MRC_HEADER = 1024;
const int MRC_LABELS = 800;
const int SWAPTRIG = 65535;

struct MrcHeader {
	int32_t nx, ny, nz, mode, nxStart, nyStart, nzStart, mx, my, mz;
	float a, b, c, alpha, beta, gamma;
	int32_t mapc, mapr, maps;
	float amin, amax, amean;
	int32_t ispg, nsymbt;
	float extra[25];
	float xOrigin, yOrigin, zOrigin;
	char map[4], machst[4];
	float arms;
	int32_t nlabl;
	char labels[MRC_LABELS];
};
static_assert(sizeof(MrcHeader) == MRC_HEADER, "MRC header is 1024 bytes");

inline void swap4(void *p) { unsigned char *b = (unsigned char *) p; std::swap(b[0], b[3]); std::swap(b[1], b[2]); }
inline void swap2(void *p) { unsigned char *b = (unsigned char *) p; std::swap(b[0], b[1]); }

inline float half_to_float(uint16_t h)
{
	const uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
	if (exp == 0)
	{
		if (man == 0) bits = sign;
		else
		{
			int e = -1;
			do { e++; man <<= 1; } while (!(man & 0x400u));
			bits = sign | (uint32_t) (127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
		}
	}
	else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
	else bits = sign | (exp + 127 - 15) << 23 | man << 13;
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

int bytes_per_pixel(int mode)
{
	switch (mode) { case 0: return 1; case 1: case 6: case 12: return 2; case 2: return 4; default: return 0; }
}

} // namespace

struct rb_mrc {
	int fd = -1;
	std::string path;
	MrcHeader h;
	bool swap = false;
	long long offset = MRC_HEADER;
	long long file_size = 0;
};

extern "C" int rb_mrc_open(const char *path, rb_mrc **out)
{
	if (!path || !out) { rb_set_error("rb_mrc_open: null argument"); return RB_ERR_ARG; }
	std::unique_ptr<rb_mrc> m(new rb_mrc);
	m->path = path;
	m->fd = open(path, O_RDONLY);
	if (m->fd < 0) { rb_set_error("rb_mrc_open: cannot open %s: %s", path, strerror(errno)); return RB_ERR_ARG; }
	struct stat st;
	if (fstat(m->fd, &st) == 0) m->file_size = (long long) st.st_size;
	if (pread(m->fd, &m->h, MRC_HEADER, 0) != MRC_HEADER)
	{
		close(m->fd);
		rb_set_error("rb_mrc_open: %s is shorter than an MRC header", path);
		return RB_ERR_ARG;
	}
	MrcHeader &h = m->h;
	if (std::abs((long long) h.mode) > SWAPTRIG || std::abs((long long) h.nx) > SWAPTRIG)       // rwMRC.h:149-158
	{
		m->swap = true;
		unsigned char *b = (unsigned char *) &h;
		for (int i = 0; i < MRC_HEADER - MRC_LABELS; i += 4) swap4(b + i);
	}
	if (h.mode == 3 || h.mode == 4)
	{
		close(m->fd);
		rb_set_error("rb_mrc_open: %s holds a transform (mode %d): only real-space images may be read", path, h.mode);
		return RB_ERR_ARG;
	}
	if (!bytes_per_pixel(h.mode))
	{
		close(m->fd);
		rb_set_error("rb_mrc_open: %s: unsupported MRC mode %d", path, h.mode);
		return RB_ERR_ARG;
	}
	if (h.nx <= 0 || h.ny <= 0 || h.nz <= 0 || h.nsymbt < 0)
	{
		close(m->fd);
		rb_set_error("rb_mrc_open: %s: invalid dimensions %d x %d x %d (nsymbt %d)", path, h.nx, h.ny, h.nz, h.nsymbt);
		return RB_ERR_ARG;
	}
	m->offset = MRC_HEADER + (long long) h.nsymbt;
	const long long need = m->offset + (long long) h.nx * h.ny * h.nz * bytes_per_pixel(h.mode);
	if (m->file_size && m->file_size < need)
	{
		close(m->fd);
		rb_set_error("rb_mrc_open: %s is truncated (%lld bytes, header promises %lld)", path, m->file_size, need);
		return RB_ERR_ARG;
	}
	*out = m.release();
	return RB_OK;
}

extern "C" int rb_mrc_info(const rb_mrc *m, int *nx, int *ny, int *nz, int *mode, float *pixel_size)
{
	if (!m) { rb_set_error("rb_mrc_info: null handle"); return RB_ERR_ARG; }
	if (nx) *nx = m->h.nx;
	if (ny) *ny = m->h.ny;
	if (nz) *nz = m->h.nz;
	if (mode) *mode = m->h.mode;
	if (pixel_size) *pixel_size = (m->h.mx && m->h.a != 0.f) ? m->h.a / (float) m->h.mx : 0.f;     // rwMRC.h:238-239
	return RB_OK;
}

// one image (0-based) of the stack -> nx*ny floats; `raw` is scratch of at least nx*ny*bytes_per_pixel
static int mrc_read_one(const rb_mrc *m, long long index, float *dst, std::vector<unsigned char> &raw)
{
	const MrcHeader &h = m->h;
	if (index < 0 || index >= h.nz)
	{
		rb_set_error("rb_mrc_read: image number %lld exceeds stack size %d of %s", index + 1, h.nz, m->path.c_str());   // rwMRC.h:176
		return RB_ERR_ARG;
	}
	const size_t npix = (size_t) h.nx * h.ny;
	const int bpp = bytes_per_pixel(h.mode);
	const size_t bytes = npix * bpp;
	unsigned char *buf = (h.mode == 2) ? (unsigned char *) dst : (raw.resize(bytes), raw.data());
	size_t got = 0;
	const long long pos = m->offset + (long long) index * (long long) bytes;
	while (got < bytes)
	{
		const ssize_t r = pread(m->fd, buf + got, bytes - got, pos + (long long) got);
		if (r < 0 && errno == EINTR) continue;
		if (r <= 0) { rb_set_error("rb_mrc_read: short read from %s (image %lld)", m->path.c_str(), index + 1); return RB_ERR_ARG; }
		got += (size_t) r;
	}
	switch (h.mode)
	{
	case 2:
		if (m->swap) for (size_t i = 0; i < npix; i++) swap4(dst + i);
		break;
	case 0:
		for (size_t i = 0; i < npix; i++) dst[i] = (float) (signed char) buf[i];
		break;
	case 1:
		for (size_t i = 0; i < npix; i++) { int16_t v; memcpy(&v, buf + 2 * i, 2); if (m->swap) swap2(&v); dst[i] = (float) v; }
		break;
	case 6:
		for (size_t i = 0; i < npix; i++) { uint16_t v; memcpy(&v, buf + 2 * i, 2); if (m->swap) swap2(&v); dst[i] = (float) v; }
		break;
	case 12:
		for (size_t i = 0; i < npix; i++) { uint16_t v; memcpy(&v, buf + 2 * i, 2); if (m->swap) swap2(&v); dst[i] = half_to_float(v); }
		break;
	}
	return RB_OK;
}

extern "C" int rb_mrc_read_images(const rb_mrc *m, const long long *indices, int count, float *dst)
{
	if (!m || (!indices && count > 0) || (!dst && count > 0)) { rb_set_error("rb_mrc_read_images: null argument"); return RB_ERR_ARG; }
	std::vector<unsigned char> raw;
	const size_t npix = (size_t) m->h.nx * m->h.ny;
	for (int i = 0; i < count; i++)
	{
		const int st = mrc_read_one(m, indices[i], dst + (size_t) i * npix, raw);
		if (st != RB_OK) return st;
	}
	return RB_OK;
}

extern "C" void rb_mrc_close(rb_mrc *m)
{
	if (!m) return;
	if (m->fd >= 0) close(m->fd);
	delete m;
}

extern "C" int rb_mrc_write(const char *path, const float *data, int nx, int ny, int nz, float pixel_size)
{
	if (!path || !data || nx <= 0 || ny <= 0 || nz <= 0) { rb_set_error("rb_mrc_write: invalid argument"); return RB_ERR_ARG; }
	MrcHeader h;
	memset(&h, 0, sizeof(h));
	h.nx = nx; h.ny = ny; h.nz = nz; h.mode = 2;
	h.mx = nx; h.my = ny; h.mz = nz;
	const float ps = pixel_size > 0.f ? pixel_size : 1.f;
	h.a = ps * nx; h.b = ps * ny; h.c = ps * nz;
	h.alpha = h.beta = h.gamma = 90.f;
	h.mapc = 1; h.mapr = 2; h.maps = 3;
	const size_t n = (size_t) nx * ny * nz;
	double mn = data[0], mx = data[0], sum = 0., sum2 = 0.;
	for (size_t i = 0; i < n; i++) { const double v = data[i]; mn = std::min(mn, v); mx = std::max(mx, v); sum += v; sum2 += v * v; }
	const double avg = sum / (double) n;
	h.amin = (float) mn; h.amax = (float) mx; h.amean = (float) avg;
	h.arms = n > 1 ? (float) std::sqrt(std::max(0., (sum2 - (double) n * avg * avg) / (double) (n - 1))) : 0.f;
	memcpy(h.map, "MAP ", 4);
	h.machst[0] = 0x44; h.machst[1] = 0x41; h.machst[2] = 0; h.machst[3] = 0;            // little-endian IEEE
	h.nlabl = 1;
	snprintf(h.labels, 80, "relion_b200 image stack");
	FILE *f = fopen(path, "wb");
	if (!f) { rb_set_error("rb_mrc_write: cannot create %s: %s", path, strerror(errno)); return RB_ERR_ARG; }
	const bool ok = fwrite(&h, MRC_HEADER, 1, f) == 1 && fwrite(data, sizeof(float), n, f) == n;
	if (fclose(f) != 0 || !ok) { rb_set_error("rb_mrc_write: write to %s failed", path); return RB_ERR_ARG; }
	return RB_OK;
}

// ---------------------------------------------------------------------------------------------
// the feed
// ---------------------------------------------------------------------------------------------
struct FeedJob {
	int ticket = 0, buffer = -1, n = 0;
	std::vector<std::string> paths;
	std::vector<long long> index;
	int next = 0, done = 0;            // next image to hand to a reader / images finished
	int status = RB_OK;
	std::string error;
};

struct rb_feed {
	int image_size = 0, max_particles = 0, depth = 0;
	bool pinned = false;
	std::vector<float *> buffers;
	std::vector<int> buffer_ticket;    // ticket using the buffer, -1: free
	std::mutex mu;
	std::condition_variable cv_work, cv_done;
	std::deque<std::shared_ptr<FeedJob>> queue;            // jobs with images left to hand out
	std::map<int, std::shared_ptr<FeedJob>> jobs;          // all unreleased jobs
	// open stacks, least recently used first in `lru`: at most `max_open` stay open (the reference keeps ONE stack open and
	// "only opens new stacks", src/ml_optimiser.cpp:10369-10377; particle sets span thousands of per-micrograph files, an
	// unbounded cache would run into RLIMIT_NOFILE).  Readers hold a shared_ptr, so an evicted stack closes when its last
	// reader is done with it.
	std::map<std::string, std::shared_ptr<rb_mrc>> stacks;
	std::deque<std::string> lru;
	size_t max_open = 64;
	std::mutex stack_mu;
	std::vector<std::thread> threads;
	int next_ticket = 1;
	bool stop = false;
};

static std::shared_ptr<rb_mrc> feed_stack(rb_feed *f, const std::string &path, std::string &err)
{
	{
		std::lock_guard<std::mutex> lk(f->stack_mu);
		auto it = f->stacks.find(path);
		if (it != f->stacks.end())
		{
			// move to the most-recently-used end
			for (auto q = f->lru.begin(); q != f->lru.end(); ++q) if (*q == path) { f->lru.erase(q); break; }
			f->lru.push_back(path);
			return it->second;
		}
	}
	// open / validate outside every lock: submit, wait and the other readers are not held up by the file system
	rb_mrc *raw = nullptr;
	if (rb_mrc_open(path.c_str(), &raw) != RB_OK) { err = rb_last_error(); return nullptr; }
	std::shared_ptr<rb_mrc> m(raw, [](rb_mrc *x) { rb_mrc_close(x); });
	if (m->h.nx != f->image_size || m->h.ny != f->image_size)
	{
		char msg[512];
		snprintf(msg, sizeof(msg), "incorrect image size: %s holds %d x %d images, the pool expects %d", path.c_str(), m->h.nx, m->h.ny, f->image_size);
		err = msg;                                                                         // src/ml_optimiser.cpp:10382-10387
		return nullptr;
	}
	std::lock_guard<std::mutex> lk(f->stack_mu);
	auto it = f->stacks.find(path);
	if (it != f->stacks.end()) return it->second;                                         // another reader opened it meanwhile
	f->stacks[path] = m;
	f->lru.push_back(path);
	while (f->lru.size() > f->max_open)
	{
		f->stacks.erase(f->lru.front());                                                   // closes once no reader holds it
		f->lru.pop_front();
	}
	return m;
}

static void feed_worker(rb_feed *f)
{
	std::vector<unsigned char> raw;
	for (;;)
	{
		std::shared_ptr<FeedJob> job;
		int i = -1;
		{
			std::unique_lock<std::mutex> lk(f->mu);
			f->cv_work.wait(lk, [&] { return f->stop || !f->queue.empty(); });
			if (f->stop) return;
			job = f->queue.front();
			i = job->next++;
			if (job->next >= job->n) f->queue.pop_front();
		}
		std::string err;
		int st = RB_OK;
		std::shared_ptr<rb_mrc> m = feed_stack(f, job->paths[i], err);
		if (!m) st = RB_ERR_ARG;
		else
		{
			const size_t npix = (size_t) f->image_size * f->image_size;
			st = mrc_read_one(m.get(), job->index[i], f->buffers[job->buffer] + (size_t) i * npix, raw);
			if (st != RB_OK) err = rb_last_error();
		}
		{
			std::lock_guard<std::mutex> lk(f->mu);
			if (st != RB_OK && job->status == RB_OK) { job->status = st; job->error = err; }
			if (++job->done == job->n) f->cv_done.notify_all();
		}
	}
}

extern "C" int rb_feed_create(int image_size, int max_particles, int depth, int n_threads, rb_feed **out)
{
	if (!out || image_size <= 0 || max_particles <= 0 || depth <= 0 || n_threads <= 0)
	{
		rb_set_error("rb_feed_create: invalid argument");
		return RB_ERR_ARG;
	}
	std::unique_ptr<rb_feed> f(new rb_feed);
	f->image_size = image_size; f->max_particles = max_particles; f->depth = depth;
	if (const char *e = getenv("RB_FEED_MAX_OPEN")) { const long v = atol(e); if (v >= 1) f->max_open = (size_t) v; }
	const size_t bytes = (size_t) max_particles * image_size * image_size * sizeof(float);
	// page-locked so that the upload of a pool is one asynchronous copy; a machine without a CUDA device (I/O unit tests)
	// gets ordinary memory - this is I/O staging, the E-step itself has no such fallback
	int ndev = 0;
	f->pinned = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
	if (!f->pinned) cudaGetLastError();
	for (int b = 0; b < depth; b++)
	{
		float *p = nullptr;
		if (f->pinned)
		{
			if (cudaHostAlloc((void **) &p, bytes, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
		}
		else if (posix_memalign((void **) &p, 4096, bytes) != 0) p = nullptr;
		if (!p)
		{
			for (float *q : f->buffers) { if (f->pinned) cudaFreeHost(q); else free(q); }
			rb_set_error("rb_feed_create: cannot allocate %zu bytes of staging memory", bytes);
			return RB_ERR_CUDA;
		}
		f->buffers.push_back(p);
		f->buffer_ticket.push_back(-1);
	}
	rb_feed *raw = f.release();
	for (int t = 0; t < n_threads; t++) raw->threads.emplace_back(feed_worker, raw);
	*out = raw;
	return RB_OK;
}

extern "C" int rb_feed_submit(rb_feed *f, const char *const *paths, const long long *index, int n_particles, int *ticket)
{
	if (!f || !paths || !index || !ticket || n_particles <= 0) { rb_set_error("rb_feed_submit: invalid argument"); return RB_ERR_ARG; }
	if (n_particles > f->max_particles)
	{
		rb_set_error("rb_feed_submit: %d particles exceed the feed's pool size %d", n_particles, f->max_particles);
		return RB_ERR_ARG;
	}
	auto job = std::make_shared<FeedJob>();
	job->n = n_particles;
	job->paths.assign(paths, paths + n_particles);
	job->index.assign(index, index + n_particles);
	std::lock_guard<std::mutex> lk(f->mu);
	for (int b = 0; b < f->depth; b++)
		if (f->buffer_ticket[b] < 0) { job->buffer = b; break; }
	if (job->buffer < 0)
	{
		rb_set_error("rb_feed_submit: all %d staging buffers are in use (release a finished pool first)", f->depth);
		return RB_ERR_STATE;
	}
	job->ticket = f->next_ticket++;
	f->buffer_ticket[job->buffer] = job->ticket;
	f->jobs[job->ticket] = job;
	f->queue.push_back(job);
	f->cv_work.notify_all();
	*ticket = job->ticket;
	return RB_OK;
}

extern "C" int rb_feed_wait(rb_feed *f, int ticket, const float **images)
{
	if (!f || !images) { rb_set_error("rb_feed_wait: invalid argument"); return RB_ERR_ARG; }
	std::unique_lock<std::mutex> lk(f->mu);
	auto it = f->jobs.find(ticket);
	if (it == f->jobs.end()) { rb_set_error("rb_feed_wait: unknown ticket %d", ticket); return RB_ERR_STATE; }
	std::shared_ptr<FeedJob> job = it->second;
	f->cv_done.wait(lk, [&] { return job->done == job->n; });
	if (job->status != RB_OK) { rb_set_error("rb_feed: %s", job->error.c_str()); return job->status; }
	*images = f->buffers[job->buffer];
	return RB_OK;
}

extern "C" int rb_feed_release(rb_feed *f, int ticket)
{
	if (!f) { rb_set_error("rb_feed_release: null handle"); return RB_ERR_ARG; }
	std::unique_lock<std::mutex> lk(f->mu);
	auto it = f->jobs.find(ticket);
	if (it == f->jobs.end()) { rb_set_error("rb_feed_release: unknown ticket %d", ticket); return RB_ERR_STATE; }
	std::shared_ptr<FeedJob> job = it->second;
	f->cv_done.wait(lk, [&] { return job->done == job->n; });      // never hand a buffer back while readers still write to it
	f->buffer_ticket[job->buffer] = -1;
	f->jobs.erase(it);
	return RB_OK;
}

extern "C" void rb_feed_destroy(rb_feed *f)
{
	if (!f) return;
	{
		std::lock_guard<std::mutex> lk(f->mu);
		f->stop = true;
		f->cv_work.notify_all();
	}
	for (auto &t : f->threads) t.join();
	f->stacks.clear(); f->lru.clear();               // shared_ptr deleters close the files
	for (float *p : f->buffers) { if (f->pinned) cudaFreeHost(p); else free(p); }
	delete f;
}
