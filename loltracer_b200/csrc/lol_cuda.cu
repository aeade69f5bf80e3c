/*
 * lol_cuda.cu -- the thin C-ABI CUDA layer of liblolb200 (sm_100a only).
 *
 * Host side: NVRTC (sm_100a) -> cudaLibrary -> persistent launch of the
 * per-scene kernel; frame plumbing (shards, de-interleave, host read-back).
 * The analogue of link_and_encode()/render_prepare() in
 * tracing_jit_renderer.dasc:60-74,416-434: there the code lands in an mmap'ed
 * RX page, here in a CUDA module.
 *
 * Static kernels in this file: the de-interleave after a gather and the FFMA
 * microbenchmark that supplies the FP32 roofline denominator.
 *
 * There is no CPU fallback: every device entry point fails with
 * LOLB200_ENODEVICE / LOLB200_ECUDA when the GPU is not usable.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvrtc.h>
#include <nvtx3/nvToolsExt.h> /* header-only: binds to a profiler's injection library at run time, no link dependency */

#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unistd.h>
#include <vector>

#include "lol_internal.h"
#include "lol_params.h"
#include "lolb200.h"

#define CUDA_TRY(expr)                                                                   \
	do {                                                                                 \
		cudaError_t e_ = (expr);                                                         \
		if (e_ != cudaSuccess) {                                                         \
			lolb200_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
			                  __FILE__, __LINE__);                                       \
			return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)        \
			           ? LOLB200_ENODEVICE                                               \
			           : LOLB200_ECUDA;                                                  \
		}                                                                                \
	} while (0)

namespace {

/* NVTX ranges around the host-side phases (lower / NVRTC / module load / launch / gather /
 * read-back): the analogue of the perf jitdump records the reference's JIT writes
 * (jitdump.c:69-129) -- a timeline tool attributes time to the phase that spent it.  Without a
 * tool attached a range costs one predictable branch. */
struct NvtxRange {
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
	NvtxRange(const NvtxRange&) = delete;
	NvtxRange& operator=(const NvtxRange&) = delete;
};

struct DeviceGuard {
	int prev = -1;
	bool ok = false;
	explicit DeviceGuard(int dev) {
		if (cudaGetDevice(&prev) != cudaSuccess)
			prev = -1;
		ok = cudaSetDevice(dev) == cudaSuccess;
	}
	~DeviceGuard() {
		if (prev >= 0)
			cudaSetDevice(prev);
	}
};

} // namespace

#define LOL_MAX_SLABS 8

struct lolb200_renderer {
	int device = 0;
	lolb200_options opt{};
	lolb200_scene* scene = nullptr;
	std::string source;
	std::vector<char> image;
	cudaLibrary_t lib = nullptr;
	cudaKernel_t kernel = nullptr;
	int sm_count = 0;
	int blocks_per_sm = 1;
	int regs = 0, smem = 0, local = 0, max_threads = 0;
	int variant = 1;
	int threads = LOLB200_KERNEL_THREADS; /* CTA size the program was generated for */
	size_t dyn_smem = 0; /* variant 2: warp-private queues */
	cudaKernel_t resume_kernel = nullptr; /* variant 4: lol_resume */
	bool long_sdf = false;                /* the program uses the guarded forms (pick_chunk_w) */
	lol_u32* queue[LOL_MAX_SLABS] = {};    /* variant 4: continuation queues of each work-counter slot */
	size_t queue_slots[LOL_MAX_SLABS] = {}; /* records per queue (one queue per class: 1 + lights) */
	int queue_classes = 0;
	lol_u32* queue_ctl = nullptr; /* per slot 64 words: per class (pushed, next to resume), then finished CTAs */
	lol_u32* counter = nullptr; /* device: [0] next chunk, [1] finished CTAs */
	lol_u64* stats = nullptr;   /* device: 8 accumulators (options.counters) */
	/* render_host staging */
	lol_u32* frame = nullptr;
	size_t frame_pixels = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr; /* read-back of finished slabs */
	cudaStream_t slab_stream[LOL_MAX_SLABS] = {}; /* slab k's launch: they may overlap */
	cudaEvent_t slab_done[LOL_MAX_SLABS] = {};
	/* A host surface that is not CUDA-pinned memory is never registered behind its owner's
	 * back (the renderer cannot know when SDL or numpy frees it): the read-back lands in
	 * this pinned frame of our own and the calling thread copies it on. */
	lol_u32* host_stage = nullptr;
	size_t host_stage_pixels = 0;
	cudaEvent_t slab_copied[LOL_MAX_SLABS] = {}; /* slab k has arrived in host memory */
	/* One launch in flight per work-counter slot: the next launch that uses a slot from a
	 * different stream waits for this event first (the slot's chunk counter, and for slot 0
	 * the longest-first buffers, are shared state). */
	cudaEvent_t slot_done[LOL_MAX_SLABS] = {};
	void* slot_stream[LOL_MAX_SLABS] = {};
	bool slot_used[LOL_MAX_SLABS] = {};
	/* a shard enqueued by host_shard_enqueue and not yet waited for */
	struct {
		bool active = false, staged = false;
		void* pixels = nullptr;
		size_t pitch_bytes = 0, world = 1, rank = 0, slabs = 0;
		int w = 0, h = 0;
		size_t begin[LOL_MAX_SLABS + 1] = {};
	} pending;
	/* longest-first chunk order (full-shard launches of lolb200_render_device) */
	struct {
		int w = 0, h = 0, rank = -1, world = 0;
		lol_u32 chunk_w = 0, n_chunks = 0;
		lol_u32 *cost = nullptr, *cost_sorted = nullptr, *ids = nullptr, *order = nullptr;
		void* tmp = nullptr;
		size_t tmp_bytes = 0, cap = 0;
		unsigned frames = 0; /* launches with this geometry */
		bool have_order = false;
		void* stream = nullptr; /* the stream the order was produced on */
	} lpt;
	cudaEvent_t t_begin = nullptr, t_end = nullptr; /* the previous host frame's render time */
	float last_render_ms = 0.f, last_copy_ms_est = 0.f;
};

/* ------------------------------------------------------------------ NVRTC -- */

/* written under a private name and renamed, so a concurrent reader (another rank
 * preparing the same scene) never sees half a file */
static void write_cache_file(const std::string& path, const void* image, size_t n) {
	char tmp[4200];
	snprintf(tmp, sizeof tmp, "%s.%ld.tmp", path.c_str(), (long)getpid());
	if (FILE* f = fopen(tmp, "wb")) {
		const bool ok = fwrite(image, 1, n, f) == n;
		if (fclose(f) == 0 && ok)
			rename(tmp, path.c_str());
		else
			remove(tmp);
	}
}

/* NVRTC arguments for the arithmetic mode; `arch` is the --gpu-architecture value */
static std::vector<const char*> nvrtc_args(const lolb200_options& opt, const char* arch) {
	std::vector<const char*> args = {arch, "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
	if (opt.arith == LOLB200_ARITH_EXACT) {
		/* One rounding per operation, IEEE div/sqrt, denormals kept: the
		 * reference is SSE scalar code without FMA (reference Makefile:3). */
		args.push_back("--fmad=false");
		args.push_back("--prec-div=true");
		args.push_back("--prec-sqrt=true");
		args.push_back("--ftz=false");
	} else {
		args.push_back("--fmad=true");
		args.push_back("--prec-div=false");
		args.push_back("--prec-sqrt=false");
		args.push_back("--ftz=true");
	}
	return args;
}

extern "C" int lolb200_compile_ptx(const char* src, const lolb200_options* o, char** ptx, size_t* len) {
	lolb200_options opt;
	if (o)
		opt = *o;
	else
		lolb200_options_default(&opt);
	if (!src || !ptx) {
		lolb200_set_error("lolb200_compile_ptx: NULL argument");
		return LOLB200_EINVAL;
	}
	*ptx = nullptr;
	NvtxRange range("lolb200: NVRTC (PTX)");
	nvrtcProgram prog;
	nvrtcResult r = nvrtcCreateProgram(&prog, src, "lol_scene.cu", 0, nullptr, nullptr);
	if (r != NVRTC_SUCCESS) {
		lolb200_set_error("nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
		return LOLB200_ECOMPILE;
	}
	std::vector<const char*> args = nvrtc_args(opt, "--gpu-architecture=compute_100a");
	r = nvrtcCompileProgram(prog, (int)args.size(), args.data());
	size_t n = 0;
	if (r != NVRTC_SUCCESS || nvrtcGetPTXSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
		size_t log_size = 0;
		nvrtcGetProgramLogSize(prog, &log_size);
		std::string text(log_size ? log_size : 1, '\0');
		if (log_size)
			nvrtcGetProgramLog(prog, &text[0]);
		lolb200_set_error("NVRTC (PTX): %s\n%.3500s", nvrtcGetErrorString(r), text.c_str());
		nvrtcDestroyProgram(&prog);
		return LOLB200_ECOMPILE;
	}
	*ptx = (char*)malloc(n + 1);
	nvrtcGetPTX(prog, *ptx);
	(*ptx)[n] = 0;
	if (len)
		*len = strlen(*ptx);
	nvrtcDestroyProgram(&prog);
	return LOLB200_OK;
}

/* SASS listing of a compiled image.  Disassembly is a toolkit TOOL, not a library: the image goes to a
 * temporary file and cuobjdump -sass (else nvdisasm) runs on it -- from $LOLB200_DISASSEMBLER, PATH,
 * $CUDA_HOME/bin or /usr/local/cuda/bin. */
extern "C" int lolb200_disassemble(const void* image, size_t size, char** sass, size_t* len) {
	if (!image || !size || !sass) {
		lolb200_set_error("lolb200_disassemble: NULL argument");
		return LOLB200_EINVAL;
	}
	*sass = nullptr;
	char path[] = "/tmp/lolb200-XXXXXX.cubin";
	const int fd = mkstemps(path, 6);
	if (fd < 0 || write(fd, image, size) != (ssize_t)size) {
		lolb200_set_error("lolb200_disassemble: cannot write %s", path);
		if (fd >= 0) {
			close(fd);
			unlink(path);
		}
		return LOLB200_EINVAL;
	}
	close(fd);
	std::vector<std::string> tools;
	if (const char* t = getenv("LOLB200_DISASSEMBLER"))
		tools.push_back(t);
	tools.push_back("cuobjdump -sass");
	if (const char* home = getenv("CUDA_HOME"))
		tools.push_back(std::string(home) + "/bin/cuobjdump -sass");
	tools.push_back("/usr/local/cuda/bin/cuobjdump -sass");
	tools.push_back("nvdisasm -c");
	tools.push_back("/usr/local/cuda/bin/nvdisasm -c");
	std::string text;
	for (const std::string& tool : tools) {
		const std::string cmd = tool + " " + path + " 2>/dev/null";
		FILE* p = popen(cmd.c_str(), "r");
		if (!p)
			continue;
		text.clear();
		char buf[8192];
		size_t got;
		while ((got = fread(buf, 1, sizeof buf, p)) > 0)
			text.append(buf, got);
		const int rc = pclose(p);
		if (rc == 0 && text.find("lol_render") != std::string::npos)
			break;
		text.clear();
	}
	unlink(path);
	if (text.empty()) {
		lolb200_set_error("lolb200_disassemble: no working cuobjdump / nvdisasm (PATH, $CUDA_HOME/bin, "
		                  "/usr/local/cuda/bin, or set LOLB200_DISASSEMBLER)");
		return LOLB200_EINVAL;
	}
	*sass = (char*)malloc(text.size() + 1);
	memcpy(*sass, text.c_str(), text.size() + 1);
	if (len)
		*len = text.size();
	return LOLB200_OK;
}

extern "C" int lolb200_compile_cubin(const char* src, const lolb200_options* o, void** image,
                                     size_t* image_size, char** log) {
	lolb200_options opt;
	if (o)
		opt = *o;
	else
		lolb200_options_default(&opt);
	if (!src || !image || !image_size) {
		lolb200_set_error("lolb200_compile_cubin: NULL argument");
		return LOLB200_EINVAL;
	}
	*image = nullptr;
	*image_size = 0;
	if (log)
		*log = nullptr;

	/* LOLB200_DUMP_DIR=<dir>: keep the generated program as <dir>/lol-<hash>.cu and
	 * compile it under that name, so -lineinfo points at a file a profiler can
	 * import (ncu --import-source on).  The analogue of the JIT's perf map. */
	std::string prog_name = "lol_scene.cu";
	if (const char* dir = getenv("LOLB200_DUMP_DIR")) {
		unsigned long long hsh = 1469598103934665603ull;
		for (const char* c = src; *c; ++c)
			hsh = (hsh ^ (unsigned char)*c) * 1099511628211ull;
		char path[4096];
		snprintf(path, sizeof path, "%s/lol-%016llx.cu", dir, hsh);
		if (FILE* f = fopen(path, "w")) {
			fputs(src, f);
			fclose(f);
			prog_name = path;
		}
	}
	/* In-process memo: a group of N devices prepares the same program N times
	 * (lolb200_group_create) -- NVRTC runs once.  Keyed by the whole program text. */
	struct Memo {
		std::string src;
		int arith;
		std::vector<char> image;
	};
	static std::mutex memo_mu;
	static std::vector<Memo> memo; /* the last few programs */
	/* LOLB200_CACHE_DIR=<dir>: compiled programs are kept as <dir>/lol-<key>.cubin,
	 * key = two 64-bit FNV-1a hashes over the program text, the arithmetic mode, the
	 * NVRTC version and the target.  A hit skips NVRTC (0.4-1.1 s per scene): the
	 * second start of the viewer on a scene prepares in milliseconds, like the
	 * CPU JIT does.  (With LOLB200_DUMP_DIR the cache is bypassed: a profiler wants
	 * the line table to point at the dumped file.) */
	std::string cache_path;
	if (const char* dir = getenv("LOLB200_CACHE_DIR")) {
		if (*dir && !getenv("LOLB200_DUMP_DIR")) {
			int vmaj = 0, vmin = 0;
			nvrtcVersion(&vmaj, &vmin);
			char salt[96];
			snprintf(salt, sizeof salt, "|sm_100a|nvrtc %d.%d|arith %d|abi %d", vmaj, vmin, opt.arith,
			         LOLB200_ABI_VERSION);
			unsigned long long h1 = 1469598103934665603ull, h2 = 0x9e3779b97f4a7c15ull;
			for (const char* part : {src, (const char*)salt})
				for (const char* c = part; *c; ++c) {
					h1 = (h1 ^ (unsigned char)*c) * 1099511628211ull;
					h2 = (h2 ^ ((unsigned char)*c + 0x100)) * 0x100000001b3ull + (h2 >> 29);
				}
			char name[64];
			snprintf(name, sizeof name, "/lol-%016llx%016llx.cubin", h1, h2);
			cache_path = std::string(dir) + name;
			if (FILE* f = fopen(cache_path.c_str(), "rb")) {
				fseek(f, 0, SEEK_END);
				long n = ftell(f);
				fseek(f, 0, SEEK_SET);
				void* buf = n > 4 ? malloc((size_t)n) : nullptr;
				const bool ok = buf && fread(buf, 1, (size_t)n, f) == (size_t)n && !memcmp(buf, "\177ELF", 4);
				fclose(f);
				if (ok) {
					*image = buf;
					*image_size = (size_t)n;
					return LOLB200_OK;
				}
				free(buf); /* truncated or foreign file: recompile and replace it */
			}
		}
	}
	if (!getenv("LOLB200_DUMP_DIR")) {
		std::lock_guard<std::mutex> lock(memo_mu);
		for (const Memo& m : memo)
			if (m.arith == opt.arith && m.src == src) {
				*image = malloc(m.image.size());
				memcpy(*image, m.image.data(), m.image.size());
				*image_size = m.image.size();
				if (!cache_path.empty()) /* no (valid) file yet: leave one for the next process */
					write_cache_file(cache_path, *image, *image_size);
				return LOLB200_OK;
			}
	}
	nvrtcProgram prog;
	nvrtcResult r = nvrtcCreateProgram(&prog, src, prog_name.c_str(), 0, nullptr, nullptr);
	if (r != NVRTC_SUCCESS) {
		lolb200_set_error("nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
		return LOLB200_ECOMPILE;
	}
	std::vector<const char*> args = nvrtc_args(opt, "--gpu-architecture=sm_100a");
	{
		NvtxRange range("lolb200: NVRTC (sm_100a)");
		r = nvrtcCompileProgram(prog, (int)args.size(), args.data());
	}
	size_t log_size = 0;
	nvrtcGetProgramLogSize(prog, &log_size);
	std::string text(log_size ? log_size : 1, '\0');
	if (log_size)
		nvrtcGetProgramLog(prog, &text[0]);
	if (log && log_size > 1) {
		*log = (char*)malloc(log_size + 1);
		memcpy(*log, text.c_str(), log_size);
		(*log)[log_size] = 0;
	}
	if (r != NVRTC_SUCCESS) {
		lolb200_set_error("NVRTC: %s\n%.3500s", nvrtcGetErrorString(r), text.c_str());
		nvrtcDestroyProgram(&prog);
		return LOLB200_ECOMPILE;
	}
	size_t n = 0;
	r = nvrtcGetCUBINSize(prog, &n);
	if (r != NVRTC_SUCCESS || n == 0) {
		lolb200_set_error("nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
		nvrtcDestroyProgram(&prog);
		return LOLB200_ECOMPILE;
	}
	*image = malloc(n);
	nvrtcGetCUBIN(prog, (char*)*image);
	*image_size = n;
	nvrtcDestroyProgram(&prog);
	if (!getenv("LOLB200_DUMP_DIR")) {
		std::lock_guard<std::mutex> lock(memo_mu);
		if (memo.size() >= 4)
			memo.erase(memo.begin());
		memo.push_back(Memo{src, opt.arith, std::vector<char>((char*)*image, (char*)*image + n)});
	}
	if (!cache_path.empty())
		write_cache_file(cache_path, *image, n);
	return LOLB200_OK;
}

/* --------------------------------------------------------------- renderer -- */

extern "C" int lolb200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" void lolb200_renderer_destroy(lolb200_renderer* r) {
	if (!r)
		return;
	{
		DeviceGuard g(r->device);
		if (r->stream) {
			cudaStreamSynchronize(r->stream);
			cudaStreamDestroy(r->stream);
		}
		cudaFree(r->lpt.cost);
		cudaFree(r->lpt.cost_sorted);
		cudaFree(r->lpt.ids);
		cudaFree(r->lpt.order);
		cudaFree(r->lpt.tmp);
		if (r->t_begin)
			cudaEventDestroy(r->t_begin);
		if (r->t_end)
			cudaEventDestroy(r->t_end);
		if (r->copy_stream) {
			cudaStreamSynchronize(r->copy_stream);
			cudaStreamDestroy(r->copy_stream);
		}
		for (cudaEvent_t e : r->slab_done)
			if (e)
				cudaEventDestroy(e);
		for (cudaEvent_t e : r->slab_copied)
			if (e)
				cudaEventDestroy(e);
		for (cudaEvent_t e : r->slot_done)
			if (e)
				cudaEventDestroy(e);
		if (r->host_stage)
			cudaFreeHost(r->host_stage);
		for (cudaStream_t st : r->slab_stream)
			if (st)
				cudaStreamDestroy(st);
		cudaFree(r->frame);
		for (lol_u32* q : r->queue)
			cudaFree(q);
		cudaFree(r->queue_ctl);
		cudaFree(r->counter);
		cudaFree(r->stats);
		if (r->lib)
			cudaLibraryUnload(r->lib);
	}
	lolb200_scene_free(r->scene);
	delete r;
}

extern "C" int lolb200_renderer_create(const lolb200_scene* s, const lolb200_options* o,
                                       int device, lolb200_renderer** out) {
	if (!s || !out) {
		lolb200_set_error("lolb200_renderer_create: NULL argument");
		return LOLB200_EINVAL;
	}
	*out = nullptr;
	int rc = lolb200_scene_check(s);
	if (rc != LOLB200_OK)
		return rc;
	int ndev = 0;
	cudaError_t ce = cudaGetDeviceCount(&ndev);
	if (ce != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		lolb200_set_error("no CUDA device available (%s); liblolb200 has no CPU fallback",
		                  ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
		return LOLB200_ENODEVICE;
	}
	if (device < 0 || device >= ndev) {
		lolb200_set_error("device %d out of range (have %d)", device, ndev);
		return LOLB200_EINVAL;
	}

	lolb200_renderer* r = new lolb200_renderer();
	r->device = device;
	if (o)
		r->opt = *o;
	else
		lolb200_options_default(&r->opt);
	r->scene = lolb200_scene_clone(s);

	size_t len = 0;
	nvtxRangePushA("lolb200: lower scene to CUDA C");
	char* src = lolb200_lower_cuda(s, &r->opt, &len);
	nvtxRangePop();
	if (!src) {
		lolb200_renderer_destroy(r);
		return LOLB200_EINVAL;
	}
	r->source.assign(src, len);
	lolb200_free(src);

	void* image = nullptr;
	size_t image_size = 0;
	rc = lolb200_compile_cubin(r->source.c_str(), &r->opt, &image, &image_size, nullptr);
	if (rc != LOLB200_OK) {
		lolb200_renderer_destroy(r);
		return rc;
	}
	r->image.assign((char*)image, (char*)image + image_size);
	free(image);

#define CREATE_TRY(expr)                                                                 \
	do {                                                                                 \
		cudaError_t e_ = (expr);                                                         \
		if (e_ != cudaSuccess) {                                                         \
			lolb200_set_error("%s failed: %s", #expr, cudaGetErrorString(e_));           \
			lolb200_renderer_destroy(r);                                                 \
			return LOLB200_ECUDA;                                                        \
		}                                                                                \
	} while (0)

	DeviceGuard g(device);
	cudaDeviceProp prop;
	CREATE_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		lolb200_set_error("device %d is sm_%d%d; this build targets sm_100a (B200) only", device,
		                  prop.major, prop.minor);
		lolb200_renderer_destroy(r);
		return LOLB200_ENODEVICE;
	}
	r->sm_count = prop.multiProcessorCount;
	NvtxRange load_range("lolb200: load module");
	CREATE_TRY(cudaLibraryLoadData(&r->lib, r->image.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
	CREATE_TRY(cudaLibraryGetKernel(&r->kernel, r->lib, "lol_render"));
	if (strstr(r->source.c_str(), "#define LOL_VARIANT 4")) {
		CREATE_TRY(cudaLibraryGetKernel(&r->resume_kernel, r->lib, "lol_resume"));
		r->queue_classes = 1 + (int)s->n_lights;
		CREATE_TRY(cudaMalloc(&r->queue_ctl, 64 * LOL_MAX_SLABS * sizeof(lol_u32)));
		CREATE_TRY(cudaMemset(r->queue_ctl, 0, 64 * LOL_MAX_SLABS * sizeof(lol_u32)));
	}
	if (strstr(r->source.c_str(), "extern \"C\" __global__ void lol_grid_build()") &&
	    !strstr(r->source.c_str(), "#define LOL_NEAR_GRID (0 &&") && !strstr(r->source.c_str(), "#define LOL_GRID_OK 0")) {
		/* the candidate grid of a pruned table loop (lol_lower.c: lol_near_grid_text): one thread per cell, once */
		cudaKernel_t build = nullptr;
		CREATE_TRY(cudaLibraryGetKernel(&build, r->lib, "lol_grid_build"));
		const char* gn = strstr(r->source.c_str(), "#define LOL_GRID_N ");
		const int n = gn ? atoi(gn + strlen("#define LOL_GRID_N ")) : 32;
		CREATE_TRY(cudaLaunchKernel((const void*)build, dim3((unsigned)((2 * n * n * n + 127) / 128)), dim3(128), nullptr, 0, nullptr));
		CREATE_TRY(cudaDeviceSynchronize());
	}
	{
		/* the lowering states what it generated */
		r->long_sdf = strstr(r->source.c_str(), "#define LOL_GUARDED 1") != nullptr;
		const char* v = strstr(r->source.c_str(), "#define LOL_VARIANT ");
		r->variant = v ? atoi(v + strlen("#define LOL_VARIANT ")) : 1;
		const char* th = strstr(r->source.c_str(), "#define LOL_THREADS ");
		if (th)
			r->threads = atoi(th + strlen("#define LOL_THREADS "));
		const char* sm = strstr(r->source.c_str(), "#define LOL_SMEM_PER_WARP ");
		const char* tw = strstr(r->source.c_str(), "#define LOL_TAB_WORDS ");
		if (r->variant != 2 && tw && strstr(r->source.c_str(), "#define LOL_TAB_IN_SMEM 1")) {
			/* the kernel copies its loop tables into shared memory (lol_lower.c: struct tabs) */
			r->dyn_smem = (size_t)atol(tw + strlen("#define LOL_TAB_WORDS ")) * sizeof(lol_u32);
			CREATE_TRY(cudaFuncSetAttribute((const void*)r->kernel,
			                                cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                (int)r->dyn_smem));
			if (r->resume_kernel)
				CREATE_TRY(cudaFuncSetAttribute((const void*)r->resume_kernel,
				                                cudaFuncAttributeMaxDynamicSharedMemorySize,
				                                (int)r->dyn_smem));
		}
		if (r->variant == 2 && sm) {
			r->dyn_smem = (size_t)atol(sm + strlen("#define LOL_SMEM_PER_WARP ")) *
			              (r->threads / 32);
			CREATE_TRY(cudaFuncSetAttribute((const void*)r->kernel,
			                                cudaFuncAttributeMaxDynamicSharedMemorySize,
			                                (int)r->dyn_smem));
		}
	}
	cudaFuncAttributes fa;
	CREATE_TRY(cudaFuncGetAttributes(&fa, (const void*)r->kernel));
	r->regs = fa.numRegs;
	r->smem = (int)(fa.sharedSizeBytes + r->dyn_smem);
	r->local = (int)fa.localSizeBytes;
	r->max_threads = fa.maxThreadsPerBlock;
	int occ = 0;
	CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)r->kernel,
	                                                         r->threads, r->dyn_smem));
	r->blocks_per_sm = occ > 0 ? occ : 1;
	/* one (next chunk, finished CTAs) pair per slab, so slab launches may overlap */
	CREATE_TRY(cudaMalloc(&r->counter, 2 * LOL_MAX_SLABS * sizeof(lol_u32)));
	CREATE_TRY(cudaMemset(r->counter, 0, 2 * LOL_MAX_SLABS * sizeof(lol_u32)));
	CREATE_TRY(cudaMalloc(&r->stats, 8 * sizeof(lol_u64)));
	CREATE_TRY(cudaMemset(r->stats, 0, 8 * sizeof(lol_u64)));
	CREATE_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
	CREATE_TRY(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
	for (cudaEvent_t& e : r->slab_done)
		CREATE_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (cudaEvent_t& e : r->slab_copied)
		CREATE_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (cudaEvent_t& e : r->slot_done)
		CREATE_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreate(&r->t_begin));
	CREATE_TRY(cudaEventCreate(&r->t_end));
	{
		int lo = 0, hi = 0; /* earlier slabs get the higher priority: they must finish first */
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		for (int k = 0; k < LOL_MAX_SLABS; ++k) {
			int prio = hi + k;
			if (prio > lo)
				prio = lo;
			CREATE_TRY(cudaStreamCreateWithPriority(&r->slab_stream[k], cudaStreamNonBlocking, prio));
		}
	}
#undef CREATE_TRY
	*out = r;
	return LOLB200_OK;
}

extern "C" const char* lolb200_renderer_source(const lolb200_renderer* r) {
	return r ? r->source.c_str() : nullptr;
}

extern "C" const void* lolb200_renderer_image(const lolb200_renderer* r, size_t* size) {
	if (!r)
		return nullptr;
	if (size)
		*size = r->image.size();
	return r->image.data();
}

extern "C" int lolb200_renderer_kernel_info(const lolb200_renderer* r, int* regs, int* smem,
                                            int* local, int* max_threads) {
	if (!r)
		return LOLB200_EINVAL;
	if (regs) *regs = r->regs;
	if (smem) *smem = r->smem;
	if (local) *local = r->local;
	if (max_threads) *max_threads = r->max_threads;
	return LOLB200_OK;
}

extern "C" size_t lolb200_shard_pixels(int w, int h, int world, int band_rows) {
	if (w <= 0 || h <= 0 || world <= 0)
		return 0;
	if (band_rows == 0)
		band_rows = LOL_BAND_ROWS;
	size_t bands = ((size_t)h + band_rows - 1) / band_rows;
	size_t local = (bands + world - 1) / world;
	return local * band_rows * (size_t)w;
}

/* Work chunks are chunk_w x 4 pixels.  Wide chunks mean fewer atomics on the
 * one counter; narrow ones keep every warp busy when a shard is small. */
static lol_u32 pick_chunk_w(const lolb200_renderer* r, int w, size_t local_bands) {
	const size_t warps = (size_t)r->sm_count * r->blocks_per_sm * (r->threads / 32);
	/* variant 3: a warp step covers 16 x 4 pixels (two per lane) */
	const lol_u32 min_w = r->variant == 3 ? 16 : 8;
	static const int forced = [] { /* LOLB200_CHUNK_W=<8|16|32|64> forces a width (A/B) */
		const char* e = getenv("LOLB200_CHUNK_W");
		return e ? atoi(e) : 0;
	}();
	if (forced >= (int)min_w && forced <= 64 && (forced & (forced - 1)) == 0)
		return (lol_u32)forced;
	/* A chunk is marched by ONE warp, tile after tile, and the frame cannot end before its longest chunk
	 * does.  Programs with the guarded forms are the ones with long distance functions (>= 3 spheres or a
	 * smooth union): they get twice as many, half as wide chunks.  Measured on B200 at 4K (chunks of 8 against
	 * 16 pixels): scene4 2.027 -> 1.999 ms (tail 63 -> 6 us), scene2 0.989 -> 0.976, scene3 0.881 -> 0.868;
	 * scene.lol (no guard, short marches) 0.644 -> 0.660: it keeps the wide ones. */
	const size_t per_warp = r->long_sdf ? 32 : 16;
	for (lol_u32 cw = r->variant == 2 ? 32 : 64; cw > min_w; cw >>= 1) {
		size_t chunks = (((size_t)w + cw - 1) / cw) * local_bands;
		if (chunks >= per_warp * warps)
			return cw;
	}
	return min_w;
}

/* How many CTAs per SM a launch uses.  A frame's time is bounded below by its slowest warp: a tile
 * with a floor-grazing ray marches ~500 distance evaluations one after the other, and with 8 warps per
 * scheduler every warp gets an eighth of the issue slots.  A whole 4K frame has 55 tiles per warp and
 * hides that; one rank's shard of eight has 7, and the slowest tile (started first: longest-first
 * order) then runs as long as the whole launch (measured: 76 us of tail in a 337 us launch).  With
 * fewer resident warps each one issues more often, so small launches trade a little throughput for
 * a shorter critical path.  LOLB200_CTAS_PER_SM=<n> forces a value (A/B). */
static int resident_ctas_per_sm(const lolb200_renderer* r, lol_u32 n_chunks) {
	static const int forced = [] {
		const char* e = getenv("LOLB200_CTAS_PER_SM");
		return e ? atoi(e) : 0;
	}();
	int ctas = r->blocks_per_sm;
	if (forced > 0)
		return forced < ctas ? forced : ctas;
	/* measured on B200, scene4 at 4K: one rank's shard of eight (6.9 tiles per warp at four CTAs per
	 * SM) 0.300 -> 0.292 ms with three CTAs per SM (tail 30 -> 5 us), two CTAs 0.325 ms; a shard of
	 * four (13.8 tiles per warp) the same either way; whole frames keep all four */
	/* round 2, with one-tile chunks and the shorter loops: a shard of four (13.7 chunks per warp) 0.5205 ->
	 * 0.5026 ms (tail 44 -> 4 us), a shard of two (27 per warp) 0.998 -> 0.990 ms, a whole frame (55 per
	 * warp) 1.931 -> 1.962 ms: programs with long distance functions switch below 32 chunks per warp */
	const size_t warps = (size_t)r->sm_count * ctas * (r->threads / 32);
	if (ctas >= 4 && (size_t)n_chunks < (r->long_sdf ? 32 : 8) * warps)
		return ctas - 1;
	return ctas;
}

__global__ void lol_iota_kernel(lol_u32* v, lol_u32 n) {
	const lol_u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		v[i] = i;
}

/* Longest-first scheduling.  The kernel records what every chunk cost (clocks); every
 * few frames the costs are sorted (cub radix sort, descending) into the order in which
 * the next frames pull their chunks.  The camera of a viewer moves slowly and a
 * benchmark repeats its frame, so last frame's costs predict this frame's.  Only for
 * launches that cover a whole shard on one stream; LOLB200_LPT=0 turns it off. */
static bool lpt_prepare(lolb200_renderer* r, const lol_params& P, void* stream) {
	static const bool enabled = [] {
		const char* e = getenv("LOLB200_LPT");
		return !(e && !strcmp(e, "0"));
	}();
	auto& L = r->lpt;
	if (!enabled || P.n_chunks < 1024)
		return false;
	if (L.w != P.w || L.h != P.h || L.rank != P.rank || L.world != P.world || L.chunk_w != P.chunk_w ||
	    L.n_chunks != P.n_chunks || L.stream != stream) {
		if (L.cap < P.n_chunks) {
			cudaFree(L.cost);
			cudaFree(L.cost_sorted);
			cudaFree(L.ids);
			cudaFree(L.order);
			cudaFree(L.tmp);
			L.cost = L.cost_sorted = L.ids = L.order = nullptr;
			L.tmp = nullptr;
			L.cap = 0;
			const size_t bytes = (size_t)P.n_chunks * sizeof(lol_u32);
			size_t tmp_bytes = 0;
			cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, L.cost, L.cost_sorted, L.ids, L.order,
			                                          (int)P.n_chunks);
			if (cudaMalloc(&L.cost, bytes) != cudaSuccess || cudaMalloc(&L.cost_sorted, bytes) != cudaSuccess ||
			    cudaMalloc(&L.ids, bytes) != cudaSuccess || cudaMalloc(&L.order, bytes) != cudaSuccess ||
			    cudaMalloc(&L.tmp, tmp_bytes ? tmp_bytes : 1) != cudaSuccess) {
				cudaGetLastError();
				return false;
			}
			L.tmp_bytes = tmp_bytes;
			L.cap = P.n_chunks;
		}
		L.w = P.w, L.h = P.h, L.rank = P.rank, L.world = P.world;
		L.chunk_w = P.chunk_w, L.n_chunks = P.n_chunks;
		L.frames = 0;
		L.have_order = false;
		L.stream = stream;
		lol_iota_kernel<<<(P.n_chunks + 255) / 256, 256, 0, (cudaStream_t)stream>>>(L.ids, P.n_chunks);
	}
	return true;
}

/* after the render launch: turn the recorded costs into the next frames' order */
static void lpt_after_launch(lolb200_renderer* r, void* stream) {
	auto& L = r->lpt;
	const unsigned f = L.frames++;
	if (f == 0 || f % 8 == 7) { /* after the first frame, then every eighth */
		size_t tmp_bytes = L.tmp_bytes;
		cub::DeviceRadixSort::SortPairsDescending(L.tmp, tmp_bytes, L.cost, L.cost_sorted, L.ids, L.order,
		                                          (int)L.n_chunks, 0, 32, (cudaStream_t)stream);
		L.have_order = true;
		/* debugging aid: LOLB200_LPT_DUMP=<file> writes the chunk costs (clocks, u32 each) of the 16th frame */
		if (f == 15)
			if (const char* path = getenv("LOLB200_LPT_DUMP")) {
				std::vector<lol_u32> host(L.n_chunks);
				cudaStreamSynchronize((cudaStream_t)stream);
				cudaMemcpy(host.data(), L.cost, host.size() * sizeof(lol_u32), cudaMemcpyDeviceToHost);
				if (FILE* fp = fopen(path, "wb")) {
					fwrite(host.data(), sizeof(lol_u32), host.size(), fp);
					fclose(fp);
				}
			}
	}
}

/* One launch over local bands [band_begin, band_begin + band_count) of the
 * rank's share (band_count = 0: all of them). */
static int launch_bands(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                        const lolb200_pixfmt* fmt, const lolb200_shard* shard, void* dst_dev,
                        size_t pitch_px, const lolb200_aux* aux, void* stream, size_t band_begin,
                        size_t band_count, int counter_slot) {
	if (!r || !dst_dev || w <= 0 || h <= 0 || pitch_px < (size_t)w) {
		lolb200_set_error("lolb200_render_device: bad argument");
		return LOLB200_EINVAL;
	}
	lolb200_shard sh = {0, 1, LOL_BAND_ROWS, 1, nullptr, 0u, 0u};
	if (shard)
		sh = *shard;
	if (sh.band_rows == 0)
		sh.band_rows = LOL_BAND_ROWS;
	if (sh.band_rows != LOL_BAND_ROWS || sh.world < 1 || sh.rank < 0 || sh.rank >= sh.world) {
		lolb200_set_error("lolb200_render_device: bad shard (rank %d of %d, band_rows %d; "
		                  "bands are %d rows)", sh.rank, sh.world, sh.band_rows, LOL_BAND_ROWS);
		return LOLB200_EINVAL;
	}
	lolb200_pixfmt pf;
	if (fmt)
		pf = *fmt;
	else
		lolb200_pixfmt_default(&pf);

	lolb200_camera c = cam ? *cam : r->scene->camera;
	lolb200_camera_basis cb;
	lolb200_camera_basis_compute(&c, w, h, &cb);

	const size_t bands = ((size_t)h + LOL_BAND_ROWS - 1) / LOL_BAND_ROWS;
	size_t local_bands =
		bands > (size_t)sh.rank ? (bands - sh.rank + sh.world - 1) / sh.world : 0;
	if (band_begin >= local_bands)
		return LOLB200_OK;
	local_bands -= band_begin;
	if (band_count && band_count < local_bands)
		local_bands = band_count;

	lol_params P;
	memset(&P, 0, sizeof P);
	P.ox = cb.origin[0]; P.oy = cb.origin[1]; P.oz = cb.origin[2];
	P.dx = cb.dir[0];    P.dy = cb.dir[1];    P.dz = cb.dir[2];
	P.rx = cb.right[0];  P.ry = cb.right[1];  P.rz = cb.right[2];
	P.ux = cb.up[0];     P.uy = cb.up[1];     P.uz = cb.up[2];
	P.cw = cb.width;
	P.ch = cb.height;
	P.fw = (float)w;
	P.fh = (float)h;
	P.w = w;
	P.h = h;
	P.rank = sh.rank;
	P.world = sh.world;
	P.dst_full = sh.dst_full_frame ? 1 : 0;
	P.pitch = (lol_u32)pitch_px;
	P.chunk_w = pick_chunk_w(r, w, local_bands);
	P.chunks_per_band = (lol_u32)(((size_t)w + P.chunk_w - 1) / P.chunk_w);
	P.n_chunks = (lol_u32)(P.chunks_per_band * local_bands);
	P.band_begin = (lol_u32)band_begin;
	P.rshift = pf.rshift; P.gshift = pf.gshift; P.bshift = pf.bshift;
	P.rloss = pf.rloss;   P.gloss = pf.gloss;   P.bloss = pf.bloss;
	P.amask = pf.amask;
	P.counter = r->counter + 2 * counter_slot;
	P.dst = (lol_u32*)dst_dev;
	if (aux) {
		P.aux_dist = aux->dist;
		P.aux_id = aux->id;
		P.aux_primary = aux->primary_steps;
		P.aux_shadow = aux->shadow_steps;
		P.timing = (lol_u64*)aux->launch_timing;
	}
	P.done_flag = (lol_u32*)sh.done_flag;
	P.done_value = sh.done_value;
	P.stats = r->stats;

	DeviceGuard g(r->device);
	/* a whole shard in one launch: chunks are pulled longest first.  (lpt_prepare may enqueue on
	 * `stream`: the slot's previous launch, if it came from another stream, goes first.) */
	if (counter_slot == 0 && r->slot_used[0] && r->slot_stream[0] != stream) {
		CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, r->slot_done[0], 0));
		r->slot_stream[0] = stream; /* ordered now */
	}
	const bool lpt = band_begin == 0 && band_count == 0 && counter_slot == 0 && lpt_prepare(r, P, stream);
	if (lpt) {
		P.cost = r->lpt.cost;
		P.order = r->lpt.have_order ? r->lpt.order : nullptr;
	}
	const size_t warps_needed = P.n_chunks;
	size_t grid = (size_t)r->sm_count * resident_ctas_per_sm(r, P.n_chunks);
	const size_t grid_needed = (warps_needed + r->threads / 32 - 1) / (r->threads / 32);
	if (grid > grid_needed)
		grid = grid_needed;
	if (r->resume_kernel) {
		/* Variant 4: room for a quarter of the launch's pixels in the continuation queue (17 words per
		 * record, one queue per class); when it overflows, lanes finish their pixels in place.
		 * Every work-counter slot has its own queues, so slab launches do not share any. */
		const size_t px = (size_t)P.n_chunks * P.chunk_w * LOL_BAND_ROWS;
		size_t want = px / 4 < 65536 ? 65536 : px / 4;
		if (const char* e = getenv("LOLB200_DEFER_QUEUE")) /* tests: a tiny queue exercises the overflow path */
			if (atol(e) > 0)
				want = (size_t)atol(e);
		want = (want + 31) & ~(size_t)31;
		if (r->queue_slots[counter_slot] < want) {
			CUDA_TRY(cudaDeviceSynchronize()); /* no launch may still use the old queue */
			cudaFree(r->queue[counter_slot]);
			r->queue[counter_slot] = nullptr;
			r->queue_slots[counter_slot] = 0;
			CUDA_TRY(cudaMalloc(&r->queue[counter_slot], want * 16 * sizeof(lol_u32) * r->queue_classes));
			r->queue_slots[counter_slot] = want;
		}
		P.q = r->queue[counter_slot];
		P.q_cap = (lol_u32)r->queue_slots[counter_slot];
		P.q_ctl = r->queue_ctl + 64 * counter_slot;
		P.cap_primary = r->opt.defer_cap_primary > 0 ? (lol_u32)r->opt.defer_cap_primary : 48u;
		P.cap_shadow = r->opt.defer_cap_shadow > 0 ? (lol_u32)r->opt.defer_cap_shadow : 24u;
	}
	/* the slot's previous launch came from another stream: it must have left the counter
	 * (and the longest-first buffers) before this one starts */
	if (r->slot_used[counter_slot] && r->slot_stream[counter_slot] != stream)
		CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, r->slot_done[counter_slot], 0));
	void* args[] = {&P};
	if (r->resume_kernel) {
		/* the completion flag (if any) belongs to the LAST launch of the pair */
		lol_params first = P;
		first.done_flag = nullptr;
		void* first_args[] = {&first};
		CUDA_TRY(cudaLaunchKernel((const void*)r->kernel, dim3((unsigned)grid), dim3((unsigned)r->threads),
		                          first_args, r->dyn_smem, (cudaStream_t)stream));
		P.timing = nullptr;
		CUDA_TRY(cudaLaunchKernel((const void*)r->resume_kernel, dim3((unsigned)((size_t)r->sm_count * r->blocks_per_sm)),
		                          dim3((unsigned)r->threads), args, r->dyn_smem, (cudaStream_t)stream));
	} else
	CUDA_TRY(cudaLaunchKernel((const void*)r->kernel, dim3((unsigned)grid), dim3((unsigned)r->threads),
	                          args, r->dyn_smem, (cudaStream_t)stream));
	if (lpt)
		lpt_after_launch(r, stream);
	CUDA_TRY(cudaEventRecord(r->slot_done[counter_slot], (cudaStream_t)stream));
	r->slot_stream[counter_slot] = stream;
	r->slot_used[counter_slot] = true;
	return LOLB200_OK;
}

extern "C" int lolb200_render_device(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                                     const lolb200_pixfmt* fmt, const lolb200_shard* shard,
                                     void* dst_dev, size_t pitch_px, const lolb200_aux* aux,
                                     void* stream) {
	NvtxRange range("lolb200: launch lol_render");
	return launch_bands(r, cam, w, h, fmt, shard, dst_dev, pitch_px, aux, stream, 0, 0, 0);
}

/* ------------------------------------------------------------ host surfaces -- */

/* Is [pixels, pixels + bytes) page-locked CUDA host memory RIGHT NOW (cudaHostAlloc, or
 * cudaHostRegister / lolb200_surface_pin by its owner)?  Asked on every frame and never
 * cached: the renderer does not own the surface, so yesterday's answer says nothing about
 * a buffer that was freed and another that now lives at the same address.  *dev_view is
 * the device-side address for zero-copy stores (NULL when the memory is not mapped). */
static bool surface_is_pinned(void* pixels, size_t bytes, void** dev_view) {
	if (dev_view)
		*dev_view = nullptr;
	cudaPointerAttributes first, last;
	if (cudaPointerGetAttributes(&first, pixels) != cudaSuccess ||
	    cudaPointerGetAttributes(&last, (char*)pixels + (bytes ? bytes - 1 : 0)) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	if (first.type != cudaMemoryTypeHost || last.type != cudaMemoryTypeHost)
		return false;
	if (dev_view && cudaHostGetDevicePointer(dev_view, pixels, 0) != cudaSuccess) {
		cudaGetLastError();
		*dev_view = nullptr;
	}
	return true;
}

/* The owner of a surface may page-lock it so that frames are DMA-ed straight into it; the
 * range must stay allocated until lolb200_surface_unpin. */
extern "C" int lolb200_surface_pin(void* pixels, size_t bytes) {
	if (!pixels || !bytes) {
		lolb200_set_error("lolb200_surface_pin: bad argument");
		return LOLB200_EINVAL;
	}
	CUDA_TRY(cudaHostRegister(pixels, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
	return LOLB200_OK;
}

extern "C" int lolb200_surface_unpin(void* pixels) {
	CUDA_TRY(cudaHostUnregister(pixels));
	return LOLB200_OK;
}

/* the renderer's own pinned frame for surfaces that are plain pageable memory */
static int stage_reserve(lolb200_renderer* r, size_t pixels) {
	if (r->host_stage_pixels >= pixels)
		return LOLB200_OK;
	if (r->host_stage)
		cudaFreeHost(r->host_stage);
	r->host_stage = nullptr;
	r->host_stage_pixels = 0;
	CUDA_TRY(cudaHostAlloc(&r->host_stage, pixels * sizeof(lol_u32), cudaHostAllocPortable));
	r->host_stage_pixels = pixels;
	return LOLB200_OK;
}

/* rows [y0, y0 + rows) of a compact w-pixel-wide frame into a surface with its own pitch */
static void copy_rows(void* pixels, size_t pitch_bytes, size_t y0, const lol_u32* src, size_t w, size_t rows) {
	char* dst = (char*)pixels + y0 * pitch_bytes;
	if (pitch_bytes == w * 4) {
		memcpy(dst, src, rows * w * 4);
		return;
	}
	for (size_t y = 0; y < rows; ++y)
		memcpy(dst + y * pitch_bytes, src + y * w, w * 4);
}

/* Slab boundaries for the render / read-back pipeline.  Equal slabs when the copy
 * takes about as long as the render (light scenes: the copy engine must never wait
 * long for the next slab).  When rendering dominates (the previous frame rendered
 * for more than twice its copy time) the slabs taper -- weights 1 2 4 8 8 4 2 1 --
 * so that the one copy nothing hides, the last, is 3 % of the bytes instead of 12 %
 * (measured at 4K: scene4 2.44 -> 2.39 ms; scene.lol would go 0.88 -> 0.97, hence
 * the switch). */
static void slab_plan(size_t bands, size_t slabs, bool taper, size_t begin[LOL_MAX_SLABS + 1]) {
	size_t w[LOL_MAX_SLABS], sum = 0, acc = 0;
	const char* plan = getenv("LOLB200_SLAB_PLAN"); /* "uniform" / "taper": force one (A/B timing) */
	const bool uniform = plan ? !strcmp(plan, "uniform") : !taper;
	for (size_t k = 0; k < slabs; ++k) {
		const size_t e = k < slabs - 1 - k ? k : slabs - 1 - k;
		w[k] = uniform ? 1 : (size_t)1 << (e < 3 ? e : 3);
		sum += w[k];
	}
	begin[0] = 0;
	for (size_t k = 0; k < slabs; ++k) {
		acc += w[k];
		size_t b = (bands * acc + sum / 2) / sum;
		if (b <= begin[k])
			b = begin[k] + 1; /* never an empty slab */
		if (b > bands || k + 1 == slabs)
			b = bands;
		begin[k + 1] = b;
	}
}

extern "C" int lolb200_render_host(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                                   const lolb200_pixfmt* fmt, void* pixels, size_t pitch_bytes) {
	if (!r || !pixels || w <= 0 || h <= 0 || pitch_bytes < (size_t)w * 4) {
		lolb200_set_error("lolb200_render_host: bad argument");
		return LOLB200_EINVAL;
	}
	NvtxRange range("lolb200: frame to host surface (slabs: render + D2H)");
	DeviceGuard g(r->device);
	const size_t need = (size_t)w * h;
	if (r->frame_pixels < need) {
		cudaFree(r->frame);
		r->frame = nullptr;
		r->frame_pixels = 0;
		CUDA_TRY(cudaMalloc(&r->frame, need * sizeof(lol_u32)));
		r->frame_pixels = need;
	}
	void* dev_view = nullptr;
	const bool pinned = surface_is_pinned(pixels, pitch_bytes * (size_t)(h - 1) + (size_t)w * 4, &dev_view);
	const char* mode = getenv("LOLB200_HOST_MODE");
	if (mode && !strcmp(mode, "mapped") && pinned && dev_view && pitch_bytes % 4 == 0) {
		/* zero-copy: the kernel stores pixels straight into the pinned surface */
		int rc0 = lolb200_render_device(r, cam, w, h, fmt, nullptr, dev_view, pitch_bytes / 4,
		                                nullptr, r->stream);
		if (rc0 != LOLB200_OK)
			return rc0;
		CUDA_TRY(cudaStreamSynchronize(r->stream));
		return LOLB200_OK;
	}
	if (!pinned) {
		int rc0 = stage_reserve(r, need);
		if (rc0 != LOLB200_OK)
			return rc0;
	}
	/* The frame is rendered as a few slabs of bands, one launch each; while slab
	 * k+1 renders, the copy engine moves slab k into the caller's surface, so the
	 * frame costs about max(kernel, PCIe copy) plus one slab's copy instead of
	 * their sum.  LOLB200_HOST_MODE=copy renders and copies the frame whole.
	 * A pageable surface receives each slab from the renderer's pinned frame as soon
	 * as the slab has arrived there (memcpy by this thread, overlapping later slabs). */
	const size_t bands = ((size_t)h + LOL_BAND_ROWS - 1) / LOL_BAND_ROWS;
	size_t slabs = (mode && !strcmp(mode, "copy")) ? 1 : LOL_MAX_SLABS;
	if (mode && !strncmp(mode, "slabs", 5) && atoi(mode + 5) > 0)
		slabs = (size_t)atoi(mode + 5) <= LOL_MAX_SLABS ? (size_t)atoi(mode + 5) : LOL_MAX_SLABS;
	if (bands < slabs * 16)
		slabs = bands / 16 ? bands / 16 : 1;
	size_t begin[LOL_MAX_SLABS + 1];
	slab_plan(bands, slabs, r->last_render_ms > 2.f * r->last_copy_ms_est && r->last_copy_ms_est > 0.f, begin);
	size_t last_k = 0;
	CUDA_TRY(cudaEventRecord(r->t_begin, r->slab_stream[0]));
	for (size_t k = 0; k < slabs && begin[k] < bands; ++k) {
		last_k = k;
		const size_t b0 = begin[k], nb = begin[k + 1] - begin[k];
		const size_t y0 = b0 * LOL_BAND_ROWS;
		const size_t rows = (y0 + nb * LOL_BAND_ROWS <= (size_t)h) ? nb * LOL_BAND_ROWS : (size_t)h - y0;
		/* Slab launches go to separate streams with their own work counters: CTAs
		 * of slab k+1 move in as those of slab k run out of chunks, so the frame
		 * has one tail, not one per slab. */
		int rc = launch_bands(r, cam, w, h, fmt, nullptr, r->frame, (size_t)w, nullptr,
		                      r->slab_stream[k], b0, nb, (int)k);
		if (rc != LOLB200_OK)
			return rc;
		CUDA_TRY(cudaEventRecord(r->slab_done[k], r->slab_stream[k]));
		CUDA_TRY(cudaStreamWaitEvent(r->copy_stream, r->slab_done[k], 0));
		if (pinned) {
			CUDA_TRY(cudaMemcpy2DAsync((char*)pixels + y0 * pitch_bytes, pitch_bytes, r->frame + y0 * w,
			                           (size_t)w * 4, (size_t)w * 4, rows, cudaMemcpyDeviceToHost,
			                           r->copy_stream));
		} else {
			CUDA_TRY(cudaMemcpyAsync(r->host_stage + y0 * w, r->frame + y0 * w, rows * (size_t)w * 4,
			                         cudaMemcpyDeviceToHost, r->copy_stream));
			CUDA_TRY(cudaEventRecord(r->slab_copied[k], r->copy_stream));
		}
	}
	CUDA_TRY(cudaEventRecord(r->t_end, r->slab_stream[last_k]));
	if (!pinned) {
		for (size_t k = 0; k <= last_k; ++k) {
			const size_t y0 = begin[k] * LOL_BAND_ROWS;
			const size_t y1 = begin[k + 1] * LOL_BAND_ROWS < (size_t)h ? begin[k + 1] * LOL_BAND_ROWS : (size_t)h;
			CUDA_TRY(cudaEventSynchronize(r->slab_copied[k]));
			copy_rows(pixels, pitch_bytes, y0, r->host_stage + y0 * w, (size_t)w, y1 - y0);
		}
	}
	CUDA_TRY(cudaStreamSynchronize(r->copy_stream));
	if (cudaEventSynchronize(r->t_end) == cudaSuccess &&
	    cudaEventElapsedTime(&r->last_render_ms, r->t_begin, r->t_end) == cudaSuccess)
		r->last_copy_ms_est = (float)((double)w * h * 4 / 45e9 * 1e3); /* PCIe 5 x16, pinned: ~45-55 GB/s */
	else
		cudaGetLastError();
	return LOLB200_OK;
}

/* One rank's share of a frame, straight into the (full-frame) host surface: the
 * rank renders its cyclic 4-row bands into a compact staging buffer and its own
 * copy engine writes each band to its final rows of `pixels`.  With one GPU per
 * rank every GPU uses its own PCIe link and NVLink is not involved at all -- for
 * a frame that has to end up in HOST memory (what renderer.h asks for) this
 * beats gathering on GPU 0 first and pushing 33 MB through one link.
 * enqueue: no synchronisation; wait: the rank's pixels are in host memory. */
static int host_shard_enqueue(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                              const lolb200_pixfmt* fmt, const lolb200_shard* shard, void* pixels,
                              size_t pitch_bytes, cudaEvent_t start_after) {
	if (!r || !pixels || !shard || w <= 0 || h <= 0 || pitch_bytes < (size_t)w * 4 || shard->world < 1 ||
	    shard->rank < 0 || shard->rank >= shard->world) {
		lolb200_set_error("lolb200_render_host_shard: bad argument");
		return LOLB200_EINVAL;
	}
	NvtxRange range("lolb200: enqueue shard (render + D2H of own bands)");
	DeviceGuard g(r->device);
	const size_t world = (size_t)shard->world, rank = (size_t)shard->rank;
	const size_t need = lolb200_shard_pixels(w, h, shard->world, 0);
	if (r->frame_pixels < need) {
		cudaFree(r->frame);
		r->frame = nullptr;
		r->frame_pixels = 0;
		CUDA_TRY(cudaMalloc(&r->frame, need * sizeof(lol_u32)));
		r->frame_pixels = need;
	}
	const bool pinned = surface_is_pinned(pixels, pitch_bytes * (size_t)(h - 1) + (size_t)w * 4, nullptr);
	if (!pinned) {
		int rc0 = stage_reserve(r, need);
		if (rc0 != LOLB200_OK)
			return rc0;
	}
	r->pending.active = false;
	const size_t bands = ((size_t)h + LOL_BAND_ROWS - 1) / LOL_BAND_ROWS;
	const size_t local_bands = bands > rank ? (bands - rank + world - 1) / world : 0;
	if (local_bands == 0)
		return LOLB200_OK;
	lolb200_shard sh = *shard;
	sh.band_rows = LOL_BAND_ROWS;
	sh.dst_full_frame = 0;
	/* fewer slabs as the shard shrinks: with eight ranks the copy is a quarter of
	 * the kernel and every slab costs its host thread API calls.  Measured on B200 (scene4 4K, e2e per frame):
	 * two ranks 4 slabs 1.150 ms against 1.217 with 2 and 1.191 with 8; four ranks 4 slabs 0.6385 against 0.6673
	 * with 2 and 0.6678 with 3; eight ranks 4 slabs 0.4404 against 0.4511 with 2 and 0.4460 with 3, timed inside
	 * one run (bench.py --e2e-slabs) */
	size_t slabs = world <= 1 ? LOL_MAX_SLABS : 4;
	if (const char* e = getenv("LOLB200_SHARD_SLABS")) /* A/B: forces the slab count of a shard */
		if (atoi(e) > 0 && atoi(e) <= LOL_MAX_SLABS)
			slabs = (size_t)atoi(e);
	if (local_bands < slabs * 16)
		slabs = local_bands / 16 ? local_bands / 16 : 1;
	size_t* begin = r->pending.begin;
	slab_plan(local_bands, slabs, false, begin);
	/* whole bands with rows back to back in the surface: one copy per slab */
	const bool one_copy = pitch_bytes == (size_t)w * 4 && h % LOL_BAND_ROWS == 0;
	for (size_t k = 0; k < slabs && begin[k] < local_bands; ++k) {
		const size_t b0 = begin[k], nb = begin[k + 1] - begin[k];
		if (start_after)
			CUDA_TRY(cudaStreamWaitEvent(r->slab_stream[k], start_after, 0));
		int rc = launch_bands(r, cam, w, h, fmt, &sh, r->frame, (size_t)w, nullptr, r->slab_stream[k],
		                      b0, nb, (int)k);
		if (rc != LOLB200_OK)
			return rc;
		CUDA_TRY(cudaEventRecord(r->slab_done[k], r->slab_stream[k]));
		CUDA_TRY(cudaStreamWaitEvent(r->copy_stream, r->slab_done[k], 0));
		if (!pinned) {
			/* pageable surface: the compact slab goes to our pinned frame; host_shard_wait hands
			 * the bands on */
			const size_t off = b0 * LOL_BAND_ROWS * (size_t)w;
			CUDA_TRY(cudaMemcpyAsync(r->host_stage + off, r->frame + off, nb * LOL_BAND_ROWS * (size_t)w * 4,
			                         cudaMemcpyDeviceToHost, r->copy_stream));
			CUDA_TRY(cudaEventRecord(r->slab_copied[k], r->copy_stream));
			continue;
		}
		if (one_copy) {
			const size_t y0 = (b0 * world + rank) * LOL_BAND_ROWS;
			CUDA_TRY(cudaMemcpy2DAsync((char*)pixels + y0 * pitch_bytes, world * LOL_BAND_ROWS * pitch_bytes,
			                           r->frame + b0 * LOL_BAND_ROWS * (size_t)w,
			                           (size_t)LOL_BAND_ROWS * w * 4, (size_t)LOL_BAND_ROWS * w * 4, nb,
			                           cudaMemcpyDeviceToHost, r->copy_stream));
			continue;
		}
		/* row j of every band of the slab: one strided 2-D copy (source rows are
		 * 4*w apart in the compact buffer, destination rows world*4 surface rows) */
		for (size_t j = 0; j < LOL_BAND_ROWS; ++j) {
			size_t rows = 0; /* local bands of the slab whose row j is inside the frame */
			for (size_t lb = b0 + nb; lb > b0; --lb)
				if ((((lb - 1) * world + rank) * LOL_BAND_ROWS + j) < (size_t)h) {
					rows = lb - b0;
					break;
				}
			if (!rows)
				continue;
			const size_t y0 = (b0 * world + rank) * LOL_BAND_ROWS + j;
			CUDA_TRY(cudaMemcpy2DAsync((char*)pixels + y0 * pitch_bytes, world * LOL_BAND_ROWS * pitch_bytes,
			                           r->frame + (b0 * LOL_BAND_ROWS + j) * (size_t)w,
			                           (size_t)LOL_BAND_ROWS * w * 4, (size_t)w * 4, rows,
			                           cudaMemcpyDeviceToHost, r->copy_stream));
		}
	}
	r->pending.active = true;
	r->pending.staged = !pinned;
	r->pending.pixels = pixels;
	r->pending.pitch_bytes = pitch_bytes;
	r->pending.world = world;
	r->pending.rank = rank;
	r->pending.slabs = slabs;
	r->pending.w = w;
	r->pending.h = h;
	return LOLB200_OK;
}

static int host_shard_wait(lolb200_renderer* r) {
	NvtxRange range("lolb200: wait for shard in host memory");
	DeviceGuard g(r->device);
	auto& q = r->pending;
	if (q.active && q.staged) {
		const size_t bands = ((size_t)q.h + LOL_BAND_ROWS - 1) / LOL_BAND_ROWS;
		const size_t local_bands = bands > q.rank ? (bands - q.rank + q.world - 1) / q.world : 0;
		for (size_t k = 0; k < q.slabs && q.begin[k] < local_bands; ++k) {
			CUDA_TRY(cudaEventSynchronize(r->slab_copied[k]));
			for (size_t lb = q.begin[k]; lb < q.begin[k + 1]; ++lb) {
				const size_t y0 = (lb * q.world + q.rank) * LOL_BAND_ROWS;
				const size_t rows = y0 + LOL_BAND_ROWS <= (size_t)q.h ? LOL_BAND_ROWS : (size_t)q.h - y0;
				copy_rows(q.pixels, q.pitch_bytes, y0, r->host_stage + lb * LOL_BAND_ROWS * (size_t)q.w,
				          (size_t)q.w, rows);
			}
		}
	}
	q.active = false;
	CUDA_TRY(cudaStreamSynchronize(r->copy_stream));
	return LOLB200_OK;
}

extern "C" int lolb200_render_host_shard(lolb200_renderer* r, const lolb200_camera* cam, int w, int h,
                                         const lolb200_pixfmt* fmt, const lolb200_shard* shard,
                                         void* pixels, size_t pitch_bytes) {
	int rc = host_shard_enqueue(r, cam, w, h, fmt, shard, pixels, pitch_bytes, nullptr);
	return rc != LOLB200_OK ? rc : host_shard_wait(r);
}

extern "C" int lolb200_read_counters(lolb200_renderer* r, uint64_t out[8]) {
	if (!r || !out)
		return LOLB200_EINVAL;
	DeviceGuard g(r->device);
	CUDA_TRY(cudaDeviceSynchronize());
	CUDA_TRY(cudaMemcpy(out, r->stats, 8 * sizeof(lol_u64), cudaMemcpyDeviceToHost));
	CUDA_TRY(cudaMemset(r->stats, 0, 8 * sizeof(lol_u64)));
	return LOLB200_OK;
}

/* ------------------------------------------------------------ de-interleave -- */

/* gathered: world compact shards back to back, each [local_band][4][w] padded
 * to shard_pixels.  One thread moves four pixels of one row. */
__global__ void __launch_bounds__(256)
lol_deinterleave_kernel(const lol_u32* __restrict__ gathered, lol_u32* __restrict__ frame, int w,
                        int h, int world, size_t shard_pixels, size_t pitch_px, int vec) {
	const int quads = (w + 3) >> 2;
	const size_t total = (size_t)quads * h;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
	     i += (size_t)gridDim.x * blockDim.x) {
		const int y = (int)(i / quads);
		const int x = (int)(i - (size_t)y * quads) << 2;
		const int band = y / LOL_BAND_ROWS;
		const int rank = band % world;
		const size_t lrow = (size_t)(band / world) * LOL_BAND_ROWS + (y % LOL_BAND_ROWS);
		const lol_u32* src = gathered + rank * shard_pixels + lrow * w + x;
		lol_u32* dst = frame + (size_t)y * pitch_px + x;
		if (vec && x + 4 <= w) {
			*reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
		} else {
			for (int k = 0; k < 4 && x + k < w; ++k)
				dst[k] = src[k];
		}
	}
}

extern "C" int lolb200_deinterleave_device(const void* gathered, void* frame, int w, int h,
                                           int world, int band_rows, size_t shard_pixels,
                                           size_t pitch_px, void* stream) {
	if (!gathered || !frame || w <= 0 || h <= 0 || world < 1 ||
	    (band_rows != 0 && band_rows != LOL_BAND_ROWS) || pitch_px < (size_t)w) {
		lolb200_set_error("lolb200_deinterleave_device: bad argument");
		return LOLB200_EINVAL;
	}
	const int vec = (w % 4 == 0) && (pitch_px % 4 == 0) && (shard_pixels % 4 == 0) &&
	                ((uintptr_t)gathered % 16 == 0) && ((uintptr_t)frame % 16 == 0);
	const size_t total = (size_t)((w + 3) / 4) * h;
	int sm = 148;
	int dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess)
		cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
	size_t blocks = (total + 255) / 256;
	if (blocks > (size_t)sm * 8)
		blocks = (size_t)sm * 8;
	lol_deinterleave_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
		(const lol_u32*)gathered, (lol_u32*)frame, w, h, world, shard_pixels, pitch_px, vec);
	CUDA_TRY(cudaGetLastError());
	return LOLB200_OK;
}

/* ------------------------------------------------- stream memory operations -- */

/* Driver entry points are looked up at run time (cudaGetDriverEntryPoint): the library has no
 * link-time dependency on libcuda and still loads on a box without a driver. */
typedef int (*lol_stream_value32_fn)(void* stream, unsigned long long addr, unsigned int value, unsigned int flags);

static lol_stream_value32_fn driver_entry(const char* name) {
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return (lol_stream_value32_fn)fn;
}

extern "C" int lolb200_stream_wait_value32(void* stream, void* dev_addr, uint32_t value) {
	static lol_stream_value32_fn fn = driver_entry("cuStreamWaitValue32");
	if (!fn) {
		lolb200_set_error("cuStreamWaitValue32 is not available from this driver");
		return LOLB200_ECUDA;
	}
	const int rc = fn(stream, (unsigned long long)(uintptr_t)dev_addr, value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
	if (rc != 0) {
		lolb200_set_error("cuStreamWaitValue32 failed: CUresult %d", rc);
		return LOLB200_ECUDA;
	}
	return LOLB200_OK;
}

extern "C" int lolb200_stream_write_value32(void* stream, void* dev_addr, uint32_t value) {
	static lol_stream_value32_fn fn = driver_entry("cuStreamWriteValue32");
	if (!fn) {
		lolb200_set_error("cuStreamWriteValue32 is not available from this driver");
		return LOLB200_ECUDA;
	}
	const int rc = fn(stream, (unsigned long long)(uintptr_t)dev_addr, value, 0x0 /* CU_STREAM_WRITE_VALUE_DEFAULT */);
	if (rc != 0) {
		lolb200_set_error("cuStreamWriteValue32 failed: CUresult %d", rc);
		return LOLB200_ECUDA;
	}
	return LOLB200_OK;
}

/* ---------------------------------------------------------------- CUDA IPC -- */

extern "C" int lolb200_ipc_export(void* dev_ptr, uint8_t handle[64]) {
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t h;
	CUDA_TRY(cudaIpcGetMemHandle(&h, dev_ptr));
	memcpy(handle, &h, 64);
	return LOLB200_OK;
}

extern "C" int lolb200_ipc_open(const uint8_t handle[64], void** dev_ptr) {
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return LOLB200_OK;
}

extern "C" int lolb200_ipc_close(void* dev_ptr) {
	CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
	return LOLB200_OK;
}

/* --------------------------------------------------- FP32 peak microbenchmark -- */

/* 8 independent FFMA chains per thread, register operands only. */
__global__ void __launch_bounds__(256) lol_ffma_peak_kernel(float* out, int iters, float a, float b) {
	float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
	float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
		}
	}
	float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	if (s == 123.456f)
		out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double lolb200_measure_fp32_peak(int device, int iters, double* ms_out) {
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
		cudaGetLastError();
		lolb200_set_error("lolb200_measure_fp32_peak: no such device");
		return -1.0;
	}
	DeviceGuard g(device);
	int sm = 0;
	cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
	if (iters <= 0)
		iters = 4096;
	const int blocks = sm * 8, threads = 256;
	float* out = nullptr;
	if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) {
		lolb200_set_error("lolb200_measure_fp32_peak: cudaMalloc failed");
		return -1.0;
	}
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 6; ++rep) { /* first reps warm up clocks and I-cache */
		cudaEventRecord(e0);
		lol_ffma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep >= 2 && ms < best)
			best = ms;
	}
	cudaError_t e = cudaGetLastError();
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(out);
	if (e != cudaSuccess) {
		lolb200_set_error("lolb200_measure_fp32_peak: %s", cudaGetErrorString(e));
		return -1.0;
	}
	if (ms_out)
		*ms_out = best;
	const double flop = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
	return flop / (best * 1e-3) / 1e12;
}

/* ------------------------------------------- several GPUs, one host process -- */

namespace {

/* NCCL is bound at run time: a process that already carries an NCCL (PyTorch
 * bundles one under the same soname) keeps using that copy, and the library
 * still loads on a box without NCCL as long as nobody asks for this gather. */
struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool load() {
		if (lib)
			return true;
		lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib)
			lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!lib)
			return false;
#define LOL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name))
		LOL_SYM(CommInitAll, "ncclCommInitAll");
		LOL_SYM(CommDestroy, "ncclCommDestroy");
		LOL_SYM(GroupStart, "ncclGroupStart");
		LOL_SYM(GroupEnd, "ncclGroupEnd");
		LOL_SYM(Send, "ncclSend");
		LOL_SYM(Recv, "ncclRecv");
		LOL_SYM(GetErrorString, "ncclGetErrorString");
#undef LOL_SYM
		return CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv && GetErrorString;
	}
};

NcclApi g_nccl;

} // namespace

struct lolb200_group {
	int n = 0;
	int gather = LOLB200_GATHER_NCCL;
	std::vector<int> devices;
	std::vector<lolb200_renderer*> renderers;
	std::vector<cudaStream_t> streams;
	std::vector<cudaEvent_t> done;       /* rank i finished its kernel */
	std::vector<ncclComm_t> comms;
	std::vector<lol_u32*> shard;         /* compact shard of rank i (rank 0: gathered[0]) */
	lol_u32* gathered = nullptr;         /* devices[0]: n shards back to back */
	lol_u32* frame = nullptr;            /* devices[0]: the complete frame */
	size_t shard_px = 0, frame_px = 0;
	int w = 0, h = 0;
	cudaEvent_t t0 = nullptr, t1 = nullptr;
	double last_ms = 0.0;
	lol_u32* host_stage = nullptr; /* pinned, ours: the way to a pageable surface (nccl / peer gathers) */
	size_t host_stage_px = 0;
};

#define NCCL_TRY(expr)                                                                   \
	do {                                                                                 \
		ncclResult_t r_ = (expr);                                                        \
		if (r_ != ncclSuccess) {                                                         \
			lolb200_set_error("%s failed: %s", #expr, g_nccl.GetErrorString(r_));        \
			return LOLB200_ECUDA;                                                        \
		}                                                                                \
	} while (0)

extern "C" void lolb200_group_destroy(lolb200_group* g) {
	if (!g)
		return;
	for (int i = 0; i < g->n; ++i) {
		if (i < (int)g->devices.size())
			cudaSetDevice(g->devices[i]);
		if (i < (int)g->streams.size() && g->streams[i]) {
			cudaStreamSynchronize(g->streams[i]);
			cudaStreamDestroy(g->streams[i]);
		}
		if (i < (int)g->done.size() && g->done[i])
			cudaEventDestroy(g->done[i]);
		if (i < (int)g->comms.size() && g->comms[i])
			g_nccl.CommDestroy(g->comms[i]);
		if (i > 0 && i < (int)g->shard.size())
			cudaFree(g->shard[i]);
		if (i < (int)g->renderers.size())
			lolb200_renderer_destroy(g->renderers[i]);
	}
	if (!g->devices.empty()) {
		cudaSetDevice(g->devices[0]);
		if (g->host_stage)
			cudaFreeHost(g->host_stage);
		cudaFree(g->gathered);
		cudaFree(g->frame);
		if (g->t0)
			cudaEventDestroy(g->t0);
		if (g->t1)
			cudaEventDestroy(g->t1);
	}
	delete g;
}

extern "C" int lolb200_group_create(const lolb200_scene* s, const lolb200_options* o,
                                    const int* devices, int n, int gather, lolb200_group** out) {
	if (!s || !out || n < 1 || n > 64 || (gather != LOLB200_GATHER_NCCL && gather != LOLB200_GATHER_PEER && gather != LOLB200_GATHER_HOST)) {
		lolb200_set_error("lolb200_group_create: bad argument");
		return LOLB200_EINVAL;
	}
	*out = nullptr;
	lolb200_group* g = new lolb200_group();
	g->n = n;
	g->gather = gather;
	for (int i = 0; i < n; ++i)
		g->devices.push_back(devices ? devices[i] : i);
	g->streams.assign(n, nullptr);
	g->done.assign(n, nullptr);
	g->comms.assign(n, nullptr);
	g->shard.assign(n, nullptr);
#define GROUP_FAIL(rc_)              \
	do {                             \
		lolb200_group_destroy(g);    \
		return (rc_);                \
	} while (0)
	for (int i = 0; i < n; ++i) {
		lolb200_renderer* r = nullptr;
		int rc = lolb200_renderer_create(s, o, g->devices[i], &r);
		if (rc != LOLB200_OK)
			GROUP_FAIL(rc);
		g->renderers.push_back(r);
		if (cudaSetDevice(g->devices[i]) != cudaSuccess ||
		    cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess) {
			lolb200_set_error("lolb200_group_create: stream setup failed on device %d", g->devices[i]);
			GROUP_FAIL(LOLB200_ECUDA);
		}
	}
	cudaSetDevice(g->devices[0]);
	cudaEventCreate(&g->t0);
	cudaEventCreate(&g->t1);
	if (n > 1 && gather == LOLB200_GATHER_NCCL) {
		if (!g_nccl.load()) {
			lolb200_set_error("cannot load libnccl.so.2: %s", dlerror());
			GROUP_FAIL(LOLB200_ECUDA);
		}
		ncclResult_t nr = g_nccl.CommInitAll(g->comms.data(), n, g->devices.data());
		if (nr != ncclSuccess) {
			lolb200_set_error("ncclCommInitAll failed: %s", g_nccl.GetErrorString(nr));
			GROUP_FAIL(LOLB200_ECUDA);
		}
	}
	if (n > 1 && gather == LOLB200_GATHER_PEER) {
		for (int i = 1; i < n; ++i) {
			int can = 0;
			if (g->devices[i] == g->devices[0])
				continue; /* the same device twice (tests on a one-GPU box): its own memory */
			cudaDeviceCanAccessPeer(&can, g->devices[i], g->devices[0]);
			if (!can) {
				lolb200_set_error("device %d cannot store into device %d (no peer access)",
				                  g->devices[i], g->devices[0]);
				GROUP_FAIL(LOLB200_ECUDA);
			}
			cudaSetDevice(g->devices[i]);
			cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[0], 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
				lolb200_set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
				GROUP_FAIL(LOLB200_ECUDA);
			}
			cudaGetLastError();
		}
	}
#undef GROUP_FAIL
	*out = g;
	return LOLB200_OK;
}

static int group_resize(lolb200_group* g, int w, int h) {
	if (g->w == w && g->h == h)
		return LOLB200_OK;
	const size_t shard_px = lolb200_shard_pixels(w, h, g->n, 0);
	cudaSetDevice(g->devices[0]);
	cudaFree(g->gathered);
	cudaFree(g->frame);
	g->gathered = g->frame = nullptr;
	CUDA_TRY(cudaMalloc(&g->frame, (size_t)w * h * sizeof(lol_u32)));
	if (g->gather == LOLB200_GATHER_NCCL && g->n > 1) {
		CUDA_TRY(cudaMalloc(&g->gathered, shard_px * g->n * sizeof(lol_u32)));
		g->shard[0] = g->gathered;
		for (int i = 1; i < g->n; ++i) {
			cudaSetDevice(g->devices[i]);
			cudaFree(g->shard[i]);
			g->shard[i] = nullptr;
			CUDA_TRY(cudaMalloc(&g->shard[i], shard_px * sizeof(lol_u32)));
		}
	}
	g->shard_px = shard_px;
	g->w = w;
	g->h = h;
	return LOLB200_OK;
}

extern "C" int lolb200_group_size(const lolb200_group* g) { return g ? g->n : 0; }

/* Share i of a frame = devices[i]'s bands, from launch to surf->pixels.  enqueue returns at
 * once, wait returns when the rows are in host memory.  Different shares may be driven by
 * different host threads at the same time (every share has its own renderer, streams and
 * staging); this is how main.c's worker threads each take a GPU. */
extern "C" int lolb200_group_share_enqueue(lolb200_group* g, int share, const lolb200_camera* cam, int w, int h,
                                           const lolb200_pixfmt* fmt, void* pixels, size_t pitch_bytes) {
	if (!g || share < 0 || share >= g->n || g->gather != LOLB200_GATHER_HOST) {
		lolb200_set_error("lolb200_group_share_enqueue: bad argument (shares exist for LOLB200_GATHER_HOST groups)");
		return LOLB200_EINVAL;
	}
	lolb200_shard sh = {share, g->n, 0, 0, nullptr, 0u, 0u};
	return host_shard_enqueue(g->renderers[share], cam, w, h, fmt, &sh, pixels, pitch_bytes, nullptr);
}

extern "C" int lolb200_group_share_wait(lolb200_group* g, int share) {
	if (!g || share < 0 || share >= g->n) {
		lolb200_set_error("lolb200_group_share_wait: bad argument");
		return LOLB200_EINVAL;
	}
	return host_shard_wait(g->renderers[share]);
}

extern "C" int lolb200_group_render_host(lolb200_group* g, const lolb200_camera* cam, int w, int h,
                                         const lolb200_pixfmt* fmt, void* pixels, size_t pitch_bytes) {
	if (!g || !pixels || w <= 0 || h <= 0 || pitch_bytes < (size_t)w * 4) {
		lolb200_set_error("lolb200_group_render_host: bad argument");
		return LOLB200_EINVAL;
	}
	if (g->n == 1)
		return lolb200_render_host(g->renderers[0], cam, w, h, fmt, pixels, pitch_bytes);
	cudaSetDevice(g->devices[0]);
	if (g->gather == LOLB200_GATHER_HOST) {
		/* every GPU renders its bands and copies them into the surface over its own
		 * PCIe link; nothing crosses NVLink */
		CUDA_TRY(cudaEventRecord(g->t0, g->streams[0]));
		for (int i = 0; i < g->n; ++i) {
			lolb200_shard sh = {i, g->n, 0, 0, nullptr, 0u, 0u};
			int rc = host_shard_enqueue(g->renderers[i], cam, w, h, fmt, &sh, pixels, pitch_bytes,
			                            i ? g->t0 : nullptr);
			if (rc != LOLB200_OK)
				return rc;
		}
		for (int i = 0; i < g->n; ++i) {
			int rc = host_shard_wait(g->renderers[i]);
			if (rc != LOLB200_OK)
				return rc;
		}
		cudaSetDevice(g->devices[0]);
		CUDA_TRY(cudaEventRecord(g->t1, g->streams[0]));
		CUDA_TRY(cudaStreamSynchronize(g->streams[0]));
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, g->t0, g->t1) == cudaSuccess)
			g->last_ms = ms;
		return LOLB200_OK;
	}
	int rc = group_resize(g, w, h);
	if (rc != LOLB200_OK)
		return rc;
	cudaSetDevice(g->devices[0]);
	CUDA_TRY(cudaEventRecord(g->t0, g->streams[0]));
	for (int i = 0; i < g->n; ++i) {
		lolb200_shard sh = {i, g->n, 0, g->gather == LOLB200_GATHER_PEER ? 1 : 0, nullptr, 0u, 0u};
		if (i > 0) {
			/* rank i starts after devices[0]'s t0 so the timing brackets all ranks */
			cudaSetDevice(g->devices[i]);
			CUDA_TRY(cudaStreamWaitEvent(g->streams[i], g->t0, 0));
		}
		void* dst = g->gather == LOLB200_GATHER_PEER ? (void*)g->frame : (void*)g->shard[i];
		rc = lolb200_render_device(g->renderers[i], cam, w, h, fmt, &sh, dst, (size_t)w, nullptr,
		                           g->streams[i]);
		if (rc != LOLB200_OK)
			return rc;
	}
	if (g->gather == LOLB200_GATHER_NCCL) {
		NvtxRange range("lolb200: NCCL gather + de-interleave");
		NCCL_TRY(g_nccl.GroupStart());
		for (int i = 1; i < g->n; ++i) {
			NCCL_TRY(g_nccl.Send(g->shard[i], g->shard_px, ncclInt32, 0, g->comms[i], g->streams[i]));
			NCCL_TRY(g_nccl.Recv(g->gathered + (size_t)i * g->shard_px, g->shard_px, ncclInt32, i,
			                     g->comms[0], g->streams[0]));
		}
		NCCL_TRY(g_nccl.GroupEnd());
		cudaSetDevice(g->devices[0]);
		rc = lolb200_deinterleave_device(g->gathered, g->frame, w, h, g->n, 0, g->shard_px, (size_t)w,
		                                 g->streams[0]);
		if (rc != LOLB200_OK)
			return rc;
	} else {
		/* peer stores: the frame is complete once every rank's kernel has retired */
		for (int i = 1; i < g->n; ++i) {
			cudaSetDevice(g->devices[i]);
			CUDA_TRY(cudaEventRecord(g->done[i], g->streams[i]));
		}
		cudaSetDevice(g->devices[0]);
		for (int i = 1; i < g->n; ++i)
			CUDA_TRY(cudaStreamWaitEvent(g->streams[0], g->done[i], 0));
	}
	cudaSetDevice(g->devices[0]);
	CUDA_TRY(cudaEventRecord(g->t1, g->streams[0]));
	if (surface_is_pinned(pixels, pitch_bytes * (size_t)(h - 1) + (size_t)w * 4, nullptr)) {
		CUDA_TRY(cudaMemcpy2DAsync(pixels, pitch_bytes, g->frame, (size_t)w * 4, (size_t)w * 4, (size_t)h,
		                           cudaMemcpyDeviceToHost, g->streams[0]));
		CUDA_TRY(cudaStreamSynchronize(g->streams[0]));
	} else {
		/* a pageable surface is not ours to register: through the group's own pinned frame */
		if (g->host_stage_px < (size_t)w * h) {
			if (g->host_stage)
				cudaFreeHost(g->host_stage);
			g->host_stage = nullptr;
			g->host_stage_px = 0;
			CUDA_TRY(cudaHostAlloc(&g->host_stage, (size_t)w * h * sizeof(lol_u32), cudaHostAllocPortable));
			g->host_stage_px = (size_t)w * h;
		}
		CUDA_TRY(cudaMemcpyAsync(g->host_stage, g->frame, (size_t)w * h * 4, cudaMemcpyDeviceToHost, g->streams[0]));
		CUDA_TRY(cudaStreamSynchronize(g->streams[0]));
		copy_rows(pixels, pitch_bytes, 0, g->host_stage, (size_t)w, (size_t)h);
	}
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, g->t0, g->t1) == cudaSuccess)
		g->last_ms = ms;
	return LOLB200_OK;
}

extern "C" double lolb200_group_last_frame_ms(const lolb200_group* g) { return g ? g->last_ms : 0.0; }
