// som_b200.cu — C-ABI library of the B200-native SOM-layer hot path (see include/som_b200.h).
//
// Kernels in this translation unit (all sm_100a):
//   som_gemm3x_pair_kernel / som_gemm3x_kernel <EPI_DIST|EPI_GRAD|EPI_RAW, F16>
//                           tcgen05/TMA GEMM in 3xFP16 or 3xTF32 + fused epilogue (som_gemm.cuh)
//   prep_rows_kernel        operand staging: row norms (+ F.normalize) and the hi/lo split (row-scaled fp16, or tf32)
//   bmu_init/decode         packed (key,index) <-> int64 BMU; bmu_decode_stat_kernel also leaves the batch statistic
//                           that scales the fp16 backward operand
//   loss_coeffs_*_kernel    fused loss + backward staging (R hi/lo, partial row / column sums)
//   adamw_stage_kernel      prototype AdamW step + staging of the new prototypes
//   nvls_allreduce_mean_kernel  two-shot NVLS all-reduce of the multi-GPU exchanges
//   neighbourhood_kernel    Gaussian grid weights, materialised on demand     (models/som_layer.py:144-152)
//   weighted_loss_kernel    mean(w * d) with w recomputed in registers        (models/som_layer.py:137-142)
//   loss_grad_kernel        G = g_out * w / (B K)
//   bwd_coeffs_kernel       R = G / d (masked), rank-1 coefficients           (ATen _euclidean_dist_backward)
//
// There is no CPU path and no fallback: on anything that is not compute capability 10.x every
// entry point fails with SOM_ERR_DEVICE.
#include "../../include/som_b200.h"
#include "som_gemm.cuh"

#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_bn_override{0};
std::atomic<int> g_kchunk{16};
std::atomic<int> g_debug{0};            // GemmShape::debug (diagnostic runs of tools/gpu_probe.py only)
std::atomic<int> g_cg_override{0};
std::atomic<unsigned long long*> g_dbg_times{nullptr};
std::atomic<int> g_streamk{0};          // -1 = never, 0 = cost model, 1 = whenever possible
std::atomic<int> g_loss_fast{1};        // 0 = always the generic loss kernel (diagnostics / tests)
std::atomic<int> g_pdl{1};              // 1 = launch with programmatic stream serialization (kernel prologues overlap the predecessor's tail)

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
// `mode` arguments carry the distance (bit 0: euclidean / cosine) and the operand precision (SOM_PREC_FP16X3)
inline bool mode_ok(int mode) { return (mode & ~(1 | SOM_PREC_FP16X3)) == 0; }
inline int mode_f16(int mode) { return (mode & SOM_PREC_FP16X3) ? 1 : 0; }

#define SOM_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t err__ = (call);                                                                     \
    if (err__ != cudaSuccess)                                                                       \
      return fail(SOM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));             \
  } while (0)

struct DeviceInfo { int sms = 0; int cc_major = 0; int dev = 0; bool ok = false; };

// Every kernel of the step is launched through this helper: with the programmatic-stream-serialization attribute the
// kernel may be scheduled while its predecessor in the stream drains; all kernels of this library execute
// griddepcontrol.wait before they touch global memory, so results are unchanged (som_set_pdl(0) switches it off).
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

int device_info(DeviceInfo& out) {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo cached;
  int dev = 0;
  SOM_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    // First call on this thread for this device: bind the primary context.  PyTorch's autograd worker threads
    // may reach us before any runtime call of theirs has done so, and cuTensorMapEncodeTiled (a driver entry point)
    // fails with CUDA_ERROR_INVALID_CONTEXT on a thread without a current context.
    SOM_CUDA(cudaFree(nullptr));
    DeviceInfo d;
    SOM_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    SOM_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    d.ok = true;
    d.dev = dev;
    cached = d;
    cached_dev = dev;
  }
  out = cached;
  if (out.cc_major != 10)
    return fail(SOM_ERR_DEVICE, "som_b200 needs a compute-capability 10.x GPU (B200, sm_100a); found major " +
                                    std::to_string(out.cc_major) + " - there is no fallback path");
  return SOM_OK;
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D tensor map of an operand matrix (fp32 containers holding tf32 values, or fp16 when `f16`): `inner` contiguous
// elements per row, `outer` rows, row pitch ld elements, box = one 128-byte span (32 fp32 / 64 fp16) x box_outer, zero
// fill outside the tensor.  K-major operands use the 128-byte swizzle with 16-byte atoms; so do MN-major fp16 operands
// (box = 64 mn x 64 k: the canonical SWIZZLE_128B MN-major panel); MN-major tf32 operands need the 32-byte-atom variant
// (matches UMMA SWIZZLE_128B_BASE32B).
int make_tmap(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer,
              bool mn_major = false, int f16 = 0) {
  auto enc = get_encode();
  if (!enc) return fail(SOM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int64_t ld_mask = f16 ? 7 : 3;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld & ld_mask) != 0)
    return fail(SOM_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple as row pitch");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {f16 ? 64u : 32u, static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (mn_major && !f16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SOM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return SOM_OK;
}

// 3-D view of an MN-major operand stored [Kred rows][mn columns] (row pitch ld): {columns of a panel, k, panel} with
// panels of 32 (tf32) or 64 (fp16) columns.  One TMA operation then brings `panels` panels of a tile (a 2-D map needs
// one operation per panel).  Only for mn % panel == 0: the panel dimension must be exact for out-of-bounds panels to
// be zero filled.
int make_tmap_mn3d(CUtensorMap* map, const float* ptr, int64_t mn, int64_t kred, int64_t ld, int panels, int f16 = 0) {
  auto enc = get_encode();
  if (!enc) return fail(SOM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int64_t pmn = som::gemm_panel_mn(f16), ld_mask = f16 ? 7 : 3;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld & ld_mask) != 0 || (mn % pmn) != 0)
    return fail(SOM_ERR_ARG, "3-D TMA operand must be 16-byte aligned, row pitch a 16-byte multiple, extent a whole number of panels");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(pmn), static_cast<cuuint64_t>(kred), static_cast<cuuint64_t>(mn / pmn)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * (f16 ? 2 : 4), 128u};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(pmn), static_cast<cuuint32_t>(som::gemm_bk(f16)), static_cast<cuuint32_t>(panels)};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SOM_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string(r));
  return SOM_OK;
}
std::atomic<int> g_tma3d{1};            // 0 = always one 2-D TMA operation per MN-major panel (diagnostics)

struct TileChoice { int cg; int bn; int sk_workers; int sk_split; };

constexpr int64_t SK_FLAG_WORDS = 4096;      // head of the GEMM workspace: stream-K flags, [worker][16] (<= 256 workers)

// Cost model (nanoseconds) behind the tile choice, calibrated on B200 with tools/gemm_time.py:
//  * tensor pipe: a k-block (32 deep, 12 tcgen05.mma for 3xTF32) of a 128 x bn tile per SM costs ~4.43 ns * bn at the
//    sustained TF32 rate (822 TFLOP/s chip-wide),
//  * L2 -> shared memory: each CTA pulls (128 + B rows it holds) * 256 bytes of hi+lo operands per k-block; the
//    chip delivers ~10 TB/s in total and at most ~100 GB/s into one SM,
//  * an epilogue of ~1.5 us + 0.08 us per drained column per tile, ~4 us of prologue / launch per kernel and
//    ~3 us for the exchange of stream-K partial tiles through L2.
// A CTA pair halves the B rows per SM (so 256 x 256 pair tiles are tensor-bound where 128 x 128 tiles are
// L2-bound) but needs enough tiles to keep all 74 pairs busy.  With a workspace the pair kernel can run stream-K:
// the tiles' k-blocks are spread evenly over the pairs, which removes the tile-count quantisation at the price of
// one partial-tile round trip through L2 per pair.
// mn_major: MN-major operands that arrive as 32 x 32 panels with one 2-D TMA operation each (4 KiB instead of
// 16 KiB per operation): measured, the chip then delivers ~7.8 TB/s instead of ~10 TB/s (fused backward: 1.09 / 1.24 /
// 0.93 us per k-block at bn = 192 / 256 / 128).  Operands whose extent is a multiple of 32 go through 3-D tensor maps
// (one operation per tile) and pay no penalty: fused backward 109.7 us at bn = 192, 109.3 us at bn = 256.
double per_kb_ns(int cg, int bn, double active_sms, bool mn_major = false) {
  const double gbs_per_sm = std::min(100.0, (mn_major ? 7800.0 : 10000.0) / active_sms);   // GB/s == bytes/ns
  const double t_l2 = (128.0 + static_cast<double>(bn) / cg) * 256.0 / gbs_per_sm;
  return std::max(4.43 * bn, t_l2);
}
// Stream-K workers for `units` k-block units of 256 x bn pair tiles (0 = do not use stream-K).
int64_t streamk_workers(int64_t units, int64_t slots, int bn, int64_t ws_floats, int phases = 1) {
  int64_t workers = std::min<int64_t>(slots, std::max<int64_t>(1, units / 4));
  workers = std::min<int64_t>(workers, (ws_floats - SK_FLAG_WORDS) / (phases * 256 * static_cast<int64_t>(bn)));
  workers = std::min<int64_t>(workers, SK_FLAG_WORDS / (16 * phases));
  return workers >= 2 ? workers : 0;
}
constexpr double HANDOVER_NS = 1300.0;      // one partial tile added by a tile's owner outside its epilogue loop
double streamk_cost_ns(int64_t units, int64_t workers, int64_t nkb_typ, int bn, bool mn_major) {
  const int64_t per_worker = (units + workers - 1) / workers;
  const double t_epi = 1500.0 + 80.0 * (static_cast<double>(bn) / 2);
  // the pair kernel always has two TMEM buffers: only the last drain of a worker is exposed, the others cost a
  // stall of the issuer when they outlast an accumulation chunk (charged at a quarter)
  const double segs = std::max(1.0, static_cast<double>(per_worker) / static_cast<double>(nkb_typ)) + 1.0;
  // a tile that is cut into many pieces: its owner adds the other pieces' partial tiles one after the other (measured
  // ~1.3 us each: flag, fetch, add; the first one is folded into the epilogue loop) - config 3's forward, one tile with
  // 384 k-blocks spread over 74 pairs, spent 95 of its 108 us there
  const double pieces = static_cast<double>(nkb_typ) / static_cast<double>(std::max<int64_t>(per_worker, 1));
  const double handover = HANDOVER_NS * std::max(0.0, pieces - 2.0);
  return static_cast<double>(per_worker) * per_kb_ns(2, bn, static_cast<double>(workers) * 2, mn_major) +
         t_epi + (segs - 1.0) * t_epi * 0.25 + 4000.0 + 2000.0 + handover;
}

TileChoice pick_tile(int64_t M, int64_t N, int64_t Kred, int sms, int b_mn, int64_t ws_floats, int bn_req, int f16 = 0) {
  const int forced_bn = bn_req > 0 ? bn_req : g_bn_override.load();
  const int forced_cg = g_cg_override.load();
  const int forced_sk = g_streamk.load();            // -1 = never, 0 = cost model, 1 = whenever possible
  // (a k-block costs the same in both precisions: 12 instructions and the same operand bytes; fp16 k-blocks are 64 deep)
  const int64_t nkb = (Kred + som::gemm_bk(f16) - 1) / som::gemm_bk(f16);
  const int pmn = som::gemm_panel_mn(f16);
  TileChoice best{1, 128, 0, 0};
  double best_cost = 1e300;
  // MN-major B read panel by panel (no 3-D tensor map: extent not a multiple of 32, single-CTA kernel, or switched off)
  auto consider = [&](int cg, int bn) {
    const bool panel_loads = b_mn != 0 && (cg == 1 || N % pmn != 0 || (bn / 2) % pmn != 0 || !g_tma3d.load());
    if (forced_cg && cg != forced_cg) return;
    if (forced_bn && bn != forced_bn) return;
    if (cg == 2 && b_mn && (bn / 2) % pmn != 0) return;
    const int64_t tiles = ((M + 128 * cg - 1) / (128 * cg)) * ((N + bn - 1) / bn);
    const int64_t slots = sms / cg;
    // epilogue of one tile: ~1.5 us + 0.08 us per column a warp drains (measured, instruction-latency bound); it hides
    // behind the next tile's mainloop only when TMEM holds two accumulator buffers (the pair kernel always does,
    // the single-CTA kernel when 4 * bn <= 512 columns)
    const double t_epi = 1500.0 + 80.0 * (static_cast<double>(bn) / cg);
    const bool overlap = cg == 2 || 4 * bn <= som::TMEM_COLS;
    if (forced_sk <= 0 || cg == 1) {
      const int64_t waves = (tiles + slots - 1) / slots;
      const double active_sms = static_cast<double>(tiles < slots ? tiles : slots) * cg;
      const double t_main = static_cast<double>(nkb) * per_kb_ns(cg, bn, active_sms, panel_loads);
      const double t_tile = overlap ? std::max(t_main, t_epi) : t_main + t_epi;
      const double cost = static_cast<double>(waves) * t_tile + (overlap ? std::min(t_main, t_epi) : 0.0) + 4000.0;
      if (cost < best_cost - 1e-9) { best_cost = cost; best = TileChoice{cg, bn, 0, 0}; }
    }
    if (cg == 2 && forced_sk >= 0 && ws_floats > 0) {
      const int64_t units = tiles * nkb;
      const int64_t workers = streamk_workers(units, slots, bn, ws_floats);
      if (workers >= 2 && tiles % workers != 0) {
        const double cost = streamk_cost_ns(units, workers, nkb, bn, panel_loads);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = TileChoice{2, bn, static_cast<int>(workers), 0}; }
      }
      // tile-aligned split-K: fewer tiles than pairs -> every tile cut into `split` equal pieces, one partial-tile
      // hand-over per piece and no slivers (preferred over even ranges at equal cost: it exchanges less)
      // The owner adds the other pieces' partial tiles one after the other (~1.3 us each: flag, fetch, add), so a long
      // reduction over few tiles (config 3: one 128 x 16 tile with 384 k-blocks) wants fewer, longer pieces than the
      // pairs would allow: the split is the one that minimises  piece time + (split - 1) hand-overs.
      const int64_t split_max = tiles > 0 ? std::min<int64_t>(slots / tiles, nkb / 8) : 0;
      int64_t split = 0;
      double split_cost = 1e300;
      for (int64_t sp = 2; sp <= split_max && tiles * sp <= workers; ++sp) {
        const int64_t piece = (nkb + sp - 1) / sp;
        // (two pieces: symmetric hand-over, each pair runs the epilogue on half of the columns)
        const double c = static_cast<double>(piece) * per_kb_ns(2, bn, static_cast<double>(tiles * sp) * 2, panel_loads) +
                         (sp == 2 ? 0.5 * t_epi + 1000.0 : t_epi + 2000.0) + 4000.0 + HANDOVER_NS * static_cast<double>(sp - 2);
        if (c < split_cost) { split_cost = c; split = sp; }
      }
      if (split >= 2) {
        const double cost = split_cost;
        if (cost < best_cost * 1.02) {
          best_cost = std::min(best_cost, cost);
          best = TileChoice{2, bn, static_cast<int>(tiles * split), static_cast<int>(split)};
        }
      }
    }
  };
  if (N <= 16 && !b_mn) consider(1, 16);
  for (int bn : {128, 96, 64, 32}) consider(1, bn);
  // CTA pairs: 256-row tiles - or, for a short M with a long reduction, split-K over many pairs (the second CTA of a
  // pair then holds zero-filled rows, but the reduction runs on up to 74 pairs instead of one SM)
  if (sms >= 2 && M > 128) {
    for (int bn : {256, 192, 128, 64}) consider(2, bn);
  } else if (sms >= 2 && nkb >= 64 && ws_floats > 0) {
    for (int bn : {64, 32}) consider(2, bn);
  }
  if (best_cost >= 1e300) {                       // overrides excluded everything: honour them literally
    best.cg = forced_cg ? forced_cg : 1;
    best.bn = forced_bn ? forced_bn : 128;
    best.sk_workers = 0;
    best.sk_split = 0;
  }
  return best;
}

// One GEMM as its caller states it: C[M,N] = A . B^T with the epilogue `epi` / `e`.
struct Problem {
  const float* a_hi; const float* a_lo; int64_t lda; int a_mn;
  const float* b_hi; const float* b_lo; int64_t ldb; int b_mn;
  int64_t M, N, Kred;
  som::EpiParams e;
  int f16 = 0;       // operands are fp16 matrices (passed through the float pointers), pitches in fp16 elements
};

int check_problem(const Problem& p, int passes) {
  if (p.M <= 0 || p.N <= 0 || p.Kred <= 0) return fail(SOM_ERR_ARG, "GEMM dimensions must be positive");
  if (p.M > (1ll << 30) || p.N > (1ll << 30) || p.Kred > (1ll << 30)) return fail(SOM_ERR_ARG, "GEMM dimension too large");
  if (!p.a_hi || !p.b_hi || (passes == 3 && (!p.a_lo || !p.b_lo))) return fail(SOM_ERR_ARG, "null GEMM operand");
  return SOM_OK;
}

// Tensor maps of one problem: A boxes of `a_rows` rows, B boxes of `b_rows` rows (K-major) or panels (MN-major).
int make_problem_maps(const Problem& p, int a_rows, int b_rows, CUtensorMap* a_hi, CUtensorMap* a_lo, CUtensorMap* b_hi,
                      CUtensorMap* b_lo) {
  int rc;
  const float* alo = p.a_lo ? p.a_lo : p.a_hi;
  const float* blo = p.b_lo ? p.b_lo : p.b_hi;
  const int f = p.f16, bk = som::gemm_bk(f);
  if (p.a_mn) { if ((rc = make_tmap(a_hi, p.a_hi, p.M, p.Kred, p.lda, bk, true, f))) return rc; if ((rc = make_tmap(a_lo, alo, p.M, p.Kred, p.lda, bk, true, f))) return rc; }
  else        { if ((rc = make_tmap(a_hi, p.a_hi, p.Kred, p.M, p.lda, a_rows, false, f))) return rc; if ((rc = make_tmap(a_lo, alo, p.Kred, p.M, p.lda, a_rows, false, f))) return rc; }
  if (p.b_mn) { if ((rc = make_tmap(b_hi, p.b_hi, p.N, p.Kred, p.ldb, bk, true, f))) return rc; if ((rc = make_tmap(b_lo, blo, p.N, p.Kred, p.ldb, bk, true, f))) return rc; }
  else        { if ((rc = make_tmap(b_hi, p.b_hi, p.Kred, p.N, p.ldb, b_rows, false, f))) return rc; if ((rc = make_tmap(b_lo, blo, p.Kred, p.N, p.ldb, b_rows, false, f))) return rc; }
  return SOM_OK;
}

void fill_shape(som::GemmShape& g, const Problem& p, int cg, int bn, int kchunk, int passes) {
  g = som::GemmShape{};
  g.M = static_cast<int>(p.M); g.N = static_cast<int>(p.N); g.Kred = static_cast<int>(p.Kred);
  g.bn = bn; g.a_mn = p.a_mn ? 1 : 0; g.b_mn = p.b_mn ? 1 : 0;
  g.kchunk = kchunk; g.passes = passes; g.nprob = 1; g.f16 = p.f16;
  g.debug = g_debug.load();
  g.dbg_times = g_dbg_times.load();
  g.tiles_m = static_cast<int>((p.M + som::BM * cg - 1) / (som::BM * cg));
  g.tiles_n = static_cast<int>((p.N + bn - 1) / bn);
}

constexpr int kMaxDevices = 64;

template <int EPI, bool F16>
int launch_single_t(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                    const CUtensorMap& tb_lo, const som::GemmShape& g, const som::EpiParams& e, int grid, size_t smem,
                    cudaStream_t st) {
  static std::atomic<bool> attr_set[kMaxDevices];        // the attribute is per device (and per kernel instance)
  int dev = 0;
  SOM_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(SOM_ERR_DEVICE, "device ordinal out of range");
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    SOM_CUDA(cudaFuncSetAttribute(som::som_gemm3x_kernel<EPI, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  som::SMEM_LIMIT));
    attr_set[dev].store(true, std::memory_order_release);
  }
  SOM_CUDA((launch_kernel(som::som_gemm3x_kernel<EPI, F16>, dim3(grid), dim3(som::NUM_THREADS), smem, st, ta_hi, ta_lo, tb_hi,
                          tb_lo, g, e)));
  g_launches.fetch_add(1);
  return SOM_OK;
}

// CTA pairs of the pair kernel that can be resident at the same time on the current device (per shared-memory size;
// cached).  The in-kernel hand-over of split-K / stream-K partial tiles needs every pair of the grid to be resident -
// an owner spins on flags that a not-yet-scheduled pair would have to raise - so schedules with more workers than this
// are not launched (MPS SM shares, green contexts or another resident kernel's shared memory can lower it below 74).
template <int EPI>
int resident_pairs(size_t smem, int* out) {
  static std::mutex mu;
  static std::map<std::pair<int, size_t>, int> cache;
  int dev = 0;
  SOM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({dev, smem});
  if (it == cache.end()) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * 1024); cfg.blockDim = dim3(som::NUM_THREADS_2CTA); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    // (the two precisions of a kernel have the same block size, register limit and shared-memory footprint)
    SOM_CUDA((cudaFuncSetAttribute(som::som_gemm3x_pair_kernel<EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   som::SMEM_LIMIT)));
    SOM_CUDA((cudaOccupancyMaxActiveClusters(&n, som::som_gemm3x_pair_kernel<EPI, false>, &cfg)));
    it = cache.emplace(std::make_pair(dev, smem), n).first;
  }
  *out = it->second;
  return SOM_OK;
}

template <int EPI, bool F16>
int launch_pair_t(const som::PairMaps& m0, const som::PairMaps& m1, const som::GemmShape& g0, const som::EpiParams& e0,
                  const som::GemmShape& g1, const som::EpiParams& e1, int grid, size_t smem, cudaStream_t st) {
  if (g0.sk_workers > 0) {
    int resident = 0;
    if (int rc = resident_pairs<EPI>(smem, &resident)) return rc;
    if (g0.sk_workers > resident)
      return fail(SOM_ERR_ARG, "split-K / stream-K schedule of " + std::to_string(g0.sk_workers) + " CTA pairs, but only " +
                                   std::to_string(resident) + " can be resident on this device right now (SM share too "
                                   "small?): pass a smaller sm_limit or no workspace");
  }
  static std::atomic<bool> attr_set[kMaxDevices];
  int dev = 0;
  SOM_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(SOM_ERR_DEVICE, "device ordinal out of range");
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    SOM_CUDA((cudaFuncSetAttribute(som::som_gemm3x_pair_kernel<EPI, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   som::SMEM_LIMIT)));
    attr_set[dev].store(true, std::memory_order_release);
  }
  SOM_CUDA((launch_kernel(som::som_gemm3x_pair_kernel<EPI, F16>, dim3(grid), dim3(som::NUM_THREADS_2CTA), smem, st, m0, m1, g0,
                          e0, g1, e1)));
  g_launches.fetch_add(1);
  return SOM_OK;
}

std::atomic<unsigned int> g_sk_token{0};

// The pair kernel adds all three 3xTF32 products into one accumulator, i.e. three tensor-core roundings per k-step
// where the two-accumulator kernel has one on its main chain.  Half the k-blocks per chain (8 instead of 16) keeps
// the rounding bias of an all-positive reduction at ~2e-6 relative (measured; budget 1e-5) while a chunk still
// lasts long enough (~7 us) for the epilogue warps to finish a tile epilogue under it.
int pair_kchunk(int kchunk, int passes) { return passes == 3 ? std::max(1, kchunk / 2) : kchunk; }

// CTA-pair launch of one or two GEMMs that share the tile width, the accumulation chunking and the epilogue kind.
int launch_pair(int epi, const Problem* probs, int nprob, int bn, int sk_workers, int sk_split, int kchunk, int passes,
                float* ws, int sms, cudaStream_t st, int sk_ph1 = 0) {
  if (bn % 32 != 0 || bn < 32 || bn > som::MAX_BN_2CTA)
    return fail(SOM_ERR_ARG, "pair tile width must be a multiple of 32 within the kernel's range");
  const int b_rows = bn / 2;                       // rows of the B tile held by one CTA
  const int f16 = probs[0].f16, pmn = som::gemm_panel_mn(f16);
  bool need64 = nprob > 1;
  for (int i = 0; i < nprob; ++i) {
    need64 = need64 || probs[i].b_mn;
    if (probs[i].f16 != f16) return fail(SOM_ERR_ARG, "the GEMMs of one launch must share the operand precision");
  }
  if (need64 && b_rows % pmn != 0)
    return fail(SOM_ERR_ARG, "pair tile width must be a multiple of two MN-major panels (64 tf32 / 128 fp16 columns)");
  som::PairMaps maps[2];
  som::GemmShape g[2];
  int64_t nwork = 0;
  for (int i = 0; i < nprob; ++i) {
    if (int rc = check_problem(probs[i], passes)) return rc;
    if (int rc = make_problem_maps(probs[i], som::BM, b_rows, &maps[i].a_hi, &maps[i].a_lo, &maps[i].b_hi, &maps[i].b_lo))
      return rc;
    fill_shape(g[i], probs[i], 2, bn, kchunk, passes);
    // MN-major operands with panel-exact extents: one 3-D TMA operation per tile instead of one per panel
    const Problem& pr = probs[i];
    if (g_tma3d.load() && pr.a_mn && pr.M % pmn == 0) {
      if (int rc = make_tmap_mn3d(&maps[i].a_hi, pr.a_hi, pr.M, pr.Kred, pr.lda, som::BM / pmn, f16)) return rc;
      if (int rc = make_tmap_mn3d(&maps[i].a_lo, pr.a_lo ? pr.a_lo : pr.a_hi, pr.M, pr.Kred, pr.lda, som::BM / pmn, f16)) return rc;
      g[i].a_3d = 1;
    }
    if (g_tma3d.load() && pr.b_mn && pr.N % pmn == 0 && b_rows % pmn == 0) {
      if (int rc = make_tmap_mn3d(&maps[i].b_hi, pr.b_hi, pr.N, pr.Kred, pr.ldb, b_rows / pmn, f16)) return rc;
      if (int rc = make_tmap_mn3d(&maps[i].b_lo, pr.b_lo ? pr.b_lo : pr.b_hi, pr.N, pr.Kred, pr.ldb, b_rows / pmn, f16)) return rc;
      g[i].b_3d = 1;
    }
    nwork += static_cast<int64_t>(g[i].tiles_m) * g[i].tiles_n;
  }
  if (nprob == 1) { maps[1] = maps[0]; g[1] = g[0]; }
  const size_t b_tile = static_cast<size_t>(b_rows) * 128;      // bytes: b_rows x one 128-byte k span (either precision)
  const size_t stage = 2 * som::A_TILE_BYTES + 2 * b_tile;
  const size_t fixed = 1024 /*alignment slack*/ + som::BAR_REGION_BYTES;
  int nst = static_cast<int>((som::SMEM_LIMIT - fixed) / stage);
  if (nst > som::MAX_STAGES) nst = som::MAX_STAGES;
  if (nst < 2) return fail(SOM_ERR_ARG, "tile does not fit shared memory");
  g[0].nstages = nst;
  g[0].nprob = nprob;
  g[0].sk_workers = ws ? sk_workers : 0;
  g[0].sk_split = g[0].sk_workers > 0 ? sk_split : 0;
  g[0].sk_ph1 = g[0].sk_split < 0 ? sk_ph1 : 0;
  if (g[0].sk_workers > 0) {
    g[0].sk_flags = reinterpret_cast<unsigned int*>(ws);
    g[0].sk_ws = ws + SK_FLAG_WORDS;
    g[0].sk_token = 0x80000000u | (g_sk_token.fetch_add(1) + 1u);
  }
  const size_t smem = fixed + nst * stage;
  const int64_t slots = sms / 2;
  const int grid = g[0].sk_workers > 0 ? g[0].sk_workers * 2 : static_cast<int>(nwork < slots ? nwork : slots) * 2;
  som::EpiParams e0 = probs[0].e, e1 = probs[nprob - 1].e;
  e0.dbg = e1.dbg = g_debug.load();
  if (f16) {
    switch (epi) {
      case som::EPI_RAW:  return launch_pair_t<som::EPI_RAW, true>(maps[0], maps[1], g[0], e0, g[1], e1, grid, smem, st);
      case som::EPI_DIST: return launch_pair_t<som::EPI_DIST, true>(maps[0], maps[1], g[0], e0, g[1], e1, grid, smem, st);
      case som::EPI_GRAD: return launch_pair_t<som::EPI_GRAD, true>(maps[0], maps[1], g[0], e0, g[1], e1, grid, smem, st);
    }
  } else {
    switch (epi) {
      case som::EPI_RAW:  return launch_pair_t<som::EPI_RAW, false>(maps[0], maps[1], g[0], e0, g[1], e1, grid, smem, st);
      case som::EPI_DIST: return launch_pair_t<som::EPI_DIST, false>(maps[0], maps[1], g[0], e0, g[1], e1, grid, smem, st);
      case som::EPI_GRAD: return launch_pair_t<som::EPI_GRAD, false>(maps[0], maps[1], g[0], e0, g[1], e1, grid, smem, st);
    }
  }
  return fail(SOM_ERR_ARG, "unknown epilogue");
}

// SMs a GEMM launch may occupy: all of them, or at most `sm_limit` (whole TPC pairs) when the caller runs a collective
// kernel beside it (a per-call argument: nothing process-wide is involved).
int effective_sms(int* sms, int sm_limit = 0) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  // never plan for more CTA pairs than can be resident together (an MPS SM share or a green context can make that fewer
  // than SMs / 2): the schedulers below hand partial tiles over between resident pairs
  int resident = 0;
  if (int rc = resident_pairs<som::EPI_GRAD>(som::SMEM_LIMIT, &resident)) return rc;
  if (resident >= 1 && 2 * resident < di.sms) di.sms = 2 * resident;
  if (sm_limit >= 2 && sm_limit < di.sms) di.sms = sm_limit & ~1;
  *sms = di.sms;
  return SOM_OK;
}

int launch_gemm(int epi, const float* a_hi, const float* a_lo, int64_t lda, int a_mn, const float* b_hi,
                const float* b_lo, int64_t ldb, int b_mn, int64_t M, int64_t N, int64_t Kred, int bn_req,
                int kchunk_req, int passes, const som::EpiParams& e, float* ws, int64_t ws_floats, cudaStream_t st,
                int sm_limit = 0, int f16 = 0) {
  int sms = 0;
  if (int rc = effective_sms(&sms, sm_limit)) return rc;
  if (passes != 1 && passes != 3) return fail(SOM_ERR_ARG, "passes must be 1 or 3");
  if (ws && (reinterpret_cast<uintptr_t>(ws) & 15) != 0) return fail(SOM_ERR_ARG, "workspace must be 16-byte aligned");
  Problem p{a_hi, a_lo, lda, a_mn, b_hi, b_lo, ldb, b_mn, M, N, Kred, e, f16};
  if (int rc = check_problem(p, passes)) return rc;
  TileChoice tc = pick_tile(M, N, Kred, sms, b_mn, ws ? ws_floats : 0, bn_req, f16);
  const int cg = tc.cg, bn = tc.bn;
  const int kchunk = kchunk_req > 0 ? kchunk_req : g_kchunk.load();
  if (cg == 2)
    return launch_pair(epi, &p, 1, bn, tc.sk_workers, tc.sk_split, pair_kchunk(kchunk, passes), passes, ws, sms, st);

  if (bn % 16 != 0 || bn < 16 || bn > som::MAX_BN) return fail(SOM_ERR_ARG, "tile width must be a multiple of 16 within the kernel's range");
  som::GemmShape g;
  fill_shape(g, p, 1, bn, kchunk, passes);
  const int pmn = som::gemm_panel_mn(f16);
  const size_t b_tile = g.b_mn ? static_cast<size_t>((bn + pmn - 1) / pmn) * (f16 ? som::PANEL_BYTES16 : som::PANEL_BYTES)
                               : static_cast<size_t>(bn) * 128;
  const size_t stage = 2 * som::A_TILE_BYTES + 2 * b_tile;
  const size_t fixed = 1024 /*alignment slack*/ + som::BAR_REGION_BYTES;
  int nst = static_cast<int>((som::SMEM_LIMIT - fixed) / stage);
  if (nst > som::MAX_STAGES) nst = som::MAX_STAGES;
  if (nst < 2) return fail(SOM_ERR_ARG, "tile does not fit shared memory");
  g.nstages = nst;
  const size_t smem = fixed + nst * stage;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (int rc = make_problem_maps(p, som::BM, bn, &ta_hi, &ta_lo, &tb_hi, &tb_lo)) return rc;
  const int64_t nwork = static_cast<int64_t>(g.tiles_m) * g.tiles_n;
  const int grid = static_cast<int>(nwork < sms ? nwork : sms);
  if (f16) {
    switch (epi) {
      case som::EPI_RAW:  return launch_single_t<som::EPI_RAW, true>(ta_hi, ta_lo, tb_hi, tb_lo, g, e, grid, smem, st);
      case som::EPI_DIST: return launch_single_t<som::EPI_DIST, true>(ta_hi, ta_lo, tb_hi, tb_lo, g, e, grid, smem, st);
      case som::EPI_GRAD: return launch_single_t<som::EPI_GRAD, true>(ta_hi, ta_lo, tb_hi, tb_lo, g, e, grid, smem, st);
    }
  } else {
    switch (epi) {
      case som::EPI_RAW:  return launch_single_t<som::EPI_RAW, false>(ta_hi, ta_lo, tb_hi, tb_lo, g, e, grid, smem, st);
      case som::EPI_DIST: return launch_single_t<som::EPI_DIST, false>(ta_hi, ta_lo, tb_hi, tb_lo, g, e, grid, smem, st);
      case som::EPI_GRAD: return launch_single_t<som::EPI_GRAD, false>(ta_hi, ta_lo, tb_hi, tb_lo, g, e, grid, smem, st);
    }
  }
  return fail(SOM_ERR_ARG, "unknown epilogue");
}

// ----------------------------------------------------------------------------------------------
// Elementwise / reduction kernels around the GEMMs
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// The same rounding (nearest, ties away from zero) for values that are finite or inf: on sm_100a cvt.rna.tf32.f32 is
// emulated by four instructions (an |x| < inf test, add, select, mask - cuobjdump); adding half an ulp to the magnitude
// bits and clearing the 13 low bits is the same function on all finite values and inf, in two.  NOT for NaN inputs (the
// canonical NaN 0x7fffffff would wrap to -0): used by the fused loss kernel only, where a NaN operand means a NaN
// distance and therefore a NaN loss that the caller sees.
__device__ __forceinline__ float tf32_rna_finite(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- 3xFP16 staging (SOM_PREC_FP16X3) -------------------------------------------------------------------------
// A row v is staged as v * 2^e = hi + lo (+ residual) with hi, lo fp16 and 2^e the power of two that puts the row's
// largest magnitude into [2^14, 2^15): fp16 has only 5 exponent bits, so the scale keeps hi AND (for every element
// within 2^-17 of the row maximum) lo in the normal range - there the split carries 22 bits like the tf32 split; smaller
// elements keep an absolute error of 2^-25, i.e. 2^-39 of the row maximum.  Scaling by a power of two is exact; the
// GEMM epilogues multiply the accumulators by 2^-e of the row and the column.  aux of such a staging holds
//   aux[0 .. rows)            |row|^2 or 1 / max(|row|, eps)   (as in the tf32 staging)
//   aux[rows .. 2 rows)       2^-e
//   aux[2 rows .. 3 rows)     2^e
//   aux[3 rows .. 3 rows + 4) statistics of the forward this staging belongs to (latent staging only; see
//                             bmu_decode_stat_kernel): [0] = bound on |R_unit| * 2^-e_b * 2^-g_k / inv_count
constexpr int F16_TOP_EXP = 14;
__device__ __forceinline__ float f16_row_scale(float amax) {
  // floor(log2(amax)) from the exponent field; zero / denormal / non-finite rows keep scale 1
  const int ex = static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu);
  if (ex == 0 || ex == 255) return 1.f;
  int e = F16_TOP_EXP - (ex - 127);
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return __uint_as_float(static_cast<uint32_t>(e + 127) << 23);
}
__device__ __forceinline__ void split_f16(float v, __half& h, __half& l) {
  h = __float2half_rn(v);
  l = __float2half_rn(v - __half2float(h));
}
// two values at a time: one packed conversion each way (cvt.rn.f16x2.f32) instead of scalar conversions and packing
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& h, uint32_t& l) {
  const __half2 hh = __floats2half2_rn(a, b);
  const float2 f = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(a - f.x, b - f.y);
  h = *reinterpret_cast<const uint32_t*>(&hh);
  l = *reinterpret_cast<const uint32_t*>(&ll);
}
// v (already normalised if the mode asks for it) * scale -> hi/lo halves at column i of the staged row
__device__ __forceinline__ void store_split4_f16(__half* ph, __half* pl, int i, float4 v, float scale) {
  uint2 h, l;
  split2_f16(v.x * scale, v.y * scale, h.x, l.x);
  split2_f16(v.z * scale, v.w * scale, h.y, l.y);
  *reinterpret_cast<uint2*>(ph + i) = h;
  *reinterpret_cast<uint2*>(pl + i) = l;
}

// GROUP threads cooperate on one row (GROUP = 32: warp per row, GROUP = 256: block per row).
struct PrepSet {
  const float* src; long long rows; long long ld_src;
  float* hi; float* lo; float* aux;
  long long* packed;     // optional: packed[row] = INT64_MAX (argmin identity) for every staged row
};

// Stages up to two matrices of the same width in one launch (latents and prototypes of one forward).
template <int GROUP, bool F16>
__global__ void __launch_bounds__(256)
prep_rows_kernel(const PrepSet sa, const PrepSet sb, long long blocks_a, int dim, int mode, long long ld_out) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  constexpr int ROWS_PER_BLOCK = 256 / GROUP;
  const int gi = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
  const bool second = static_cast<long long>(blockIdx.x) >= blocks_a;
  const PrepSet& ps = second ? sb : sa;
  const float* __restrict__ src = ps.src;
  float* __restrict__ hi = ps.hi;
  float* __restrict__ lo = ps.lo;
  float* __restrict__ aux = ps.aux;
  const long long rows = ps.rows, ld_src = ps.ld_src;
  const long long row = (static_cast<long long>(blockIdx.x) - (second ? blocks_a : 0)) * ROWS_PER_BLOCK + gi;
  __shared__ float red[8];
  __shared__ float redm[8];
  const bool active = row < rows;
  const float* p = src + (active ? row : 0) * ld_src;
  const bool vec = (dim & 3) == 0 && (ld_src & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;

  // rows of up to PREP_CACHE * GROUP * 4 elements are read ONCE: the values wait in registers for the row norm
  constexpr int PREP_CACHE = 8;
  const bool cached = vec && dim <= PREP_CACHE * GROUP * 4;
  float4 cache[PREP_CACHE];
  float ss = 0.f, amax = 0.f;                    // amax: largest magnitude of the row (fp16 staging only)
  if (active) {
    if (cached) {
#pragma unroll
      for (int j = 0; j < PREP_CACHE; ++j) {
        const int i = (gt + j * GROUP) * 4;
        cache[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < dim) cache[j] = __ldg(reinterpret_cast<const float4*>(p + i));
        const float4 v = cache[j];
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        if constexpr (F16) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else if (vec) {
      for (int i = gt * 4; i < dim; i += GROUP * 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        if constexpr (F16) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else {
      for (int i = gt; i < dim; i += GROUP) {
        const float v = __ldg(p + i);
        ss = fmaf(v, v, ss);
        if constexpr (F16) amax = fmaxf(amax, fabsf(v));
      }
    }
  }
  ss = warp_sum(ss);
  if constexpr (F16) amax = warp_max(amax);
  if constexpr (GROUP == 256) {
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = ss; redm[threadIdx.x >> 5] = amax; }
    __syncthreads();
    ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += red[i];
    if constexpr (F16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) amax = fmaxf(amax, redm[i]);
    }
  }
  if constexpr (F16) {
    // the forward's statistics slot starts from zero (bmu_decode_stat_kernel raises it with atomicMax later in the stream)
    if (blockIdx.x == 0 && threadIdx.x == 0 && sa.packed) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sa.aux[3 * sa.rows + i] = 0.f;
    }
  }
  if (!active) return;
  if (ps.packed && gt == 0) ps.packed[row] = 0x7fffffffffffffffLL;
  float denom = 1.f;
  if (mode == 1) {
    denom = fmaxf(sqrtf(ss), 1e-12f);            // F.normalize(p=2, eps=1e-12)
    if (gt == 0) aux[row] = 1.f / denom;
  } else if (gt == 0) {
    aux[row] = ss;
  }
  if constexpr (F16) {
    // the normalised row's largest magnitude is amax / denom (division is monotone), its scale the power of two that
    // lifts it into [2^14, 2^15)
    const float scale = f16_row_scale(mode == 1 ? amax / denom : amax);
    if (gt == 0) { aux[rows + row] = 1.f / scale; aux[2 * rows + row] = scale; }
    __half* ph16 = reinterpret_cast<__half*>(hi) + row * ld_out;
    __half* pl16 = reinterpret_cast<__half*>(lo) + row * ld_out;
    const int dim_out16 = static_cast<int>(ld_out);
    auto norm4 = [&](float4 v) {
      if (mode == 1) { v.x = v.x / denom; v.y = v.y / denom; v.z = v.z / denom; v.w = v.w / denom; }
      return v;
    };
    if (cached) {
#pragma unroll
      for (int j = 0; j < PREP_CACHE; ++j) {
        const int i = (gt + j * GROUP) * 4;
        if (i < dim_out16) store_split4_f16(ph16, pl16, i, i < dim ? norm4(cache[j]) : make_float4(0.f, 0.f, 0.f, 0.f), scale);
      }
    } else if (vec) {
      for (int i = gt * 4; i < dim_out16; i += GROUP * 4)
        store_split4_f16(ph16, pl16, i, i < dim ? norm4(__ldg(reinterpret_cast<const float4*>(p + i))) : make_float4(0.f, 0.f, 0.f, 0.f), scale);
    } else {
      for (int i = gt; i < dim_out16; i += GROUP) {
        __half h = __float2half_rn(0.f), l = h;
        if (i < dim) {
          float v = __ldg(p + i);
          if (mode == 1) v = v / denom;
          split_f16(v * scale, h, l);
        }
        ph16[i] = h;
        pl16[i] = l;
      }
    }
    return;
  }
  float* ph = hi + row * ld_out;
  float* pl = lo + row * ld_out;
  const int dim_out = static_cast<int>(ld_out);
  auto split4 = [&](float4 v, float4& h, float4& l) {
    if (mode == 1) { v.x = v.x / denom; v.y = v.y / denom; v.z = v.z / denom; v.w = v.w / denom; }
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
  };
  if (cached) {
#pragma unroll
    for (int j = 0; j < PREP_CACHE; ++j) {
      const int i = (gt + j * GROUP) * 4;
      if (i < dim_out) {
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
        if (i < dim) split4(cache[j], h, l);
        *reinterpret_cast<float4*>(ph + i) = h;
        *reinterpret_cast<float4*>(pl + i) = l;
      }
    }
  } else if (vec) {
    for (int i = gt * 4; i < dim_out; i += GROUP * 4) {
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
      if (i < dim) split4(__ldg(reinterpret_cast<const float4*>(p + i)), h, l);
      *reinterpret_cast<float4*>(ph + i) = h;
      *reinterpret_cast<float4*>(pl + i) = l;
    }
  } else {
    for (int i = gt; i < dim_out; i += GROUP) {
      float h = 0.f, l = 0.f;
      if (i < dim) {
        float v = __ldg(p + i);
        if (mode == 1) v = v / denom;
        h = tf32_rna(v);
        l = tf32_rna(v - h);
      }
      ph[i] = h;
      pl[i] = l;
    }
  }
}

__global__ void bmu_init_kernel(long long* packed, long long n) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) packed[i] = 0x7fffffffffffffffLL;
}

__global__ void bmu_decode_kernel(const long long* __restrict__ packed, long long n, long long k_total,
                                  long long* __restrict__ bmu, float* __restrict__ min_key) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long p = packed[i];
  long long idx = p & 0xffffffffLL;
  if (idx >= k_total) idx = 0;                    // NaN / all-inf rows: stay in bounds
  bmu[i] = idx;
  if (min_key) {
    int k = static_cast<int>(p >> 32);
    k ^= (k >> 31) & 0x7fffffff;
    min_key[i] = __int_as_float(k);
  }
}

// 3xFP16: BMU decode plus the statistic that lets the loss kernel place the staged backward operand
//   R^[b,k] = R_unit[b,k] * 2^-e_b * 2^-g_k * S          (R_unit = inv_count * w / d, or inv_count * w for cosine)
// into the fp16 range with ONE power of two S for the whole batch (the gradient GEMM over b cannot absorb a per-row
// factor, the one over k no per-column factor; 2^-e_b and 2^-g_k cancel the scales of the staged latents / prototypes).
// Since w <= 1 and d >= d_bmu(b):  R^ / (inv_count * S) <= 2^-e_b / d_bmu(b) * max_k 2^-g_k.  Every block takes the
// maximum of the prototype scales itself (K floats from L2), the maximum over the rows is combined with atomicMax on the
// bit pattern of the (positive) float - order independent, hence deterministic; the slot was zeroed by this forward's
// staging kernel.  Rows whose BMU distance is zero (R is masked there) or below 2^-20 |x| (pure cancellation noise in
// fp32) do not take part; the loss kernel saturates what would leave the fp16 range.
__global__ void __launch_bounds__(256)
bmu_decode_stat_kernel(const long long* __restrict__ packed, long long n, long long k_total, long long* __restrict__ bmu,
                       float* __restrict__ min_key, float* __restrict__ x_aux, const float* __restrict__ w_aux,
                       long long K, int dmode) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  __shared__ float redv[8], redw[8];
  float mw = 0.f;
  for (long long k = threadIdx.x; k < K; k += 256) mw = fmaxf(mw, __ldg(w_aux + K + k));
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (i < n) {
    const long long p = packed[i];
    long long idx = p & 0xffffffffLL;
    if (idx >= k_total) idx = 0;                    // NaN / all-inf rows: stay in bounds
    if (bmu) bmu[i] = idx;
    int kb = static_cast<int>(p >> 32);
    kb ^= (kb >> 31) & 0x7fffffff;
    const float key = __int_as_float(kb);          // squared BMU distance (euclidean) or BMU distance (cosine)
    if (min_key) min_key[i] = key;
    const float xs = x_aux[n + i];
    if (dmode == 1) val = xs;
    else if (key > 0.f && key >= 9.094947e-13f * x_aux[i]) val = xs * rsqrtf(key);     // d >= 2^-20 |x|
    if (!(val < 3.0e38f)) val = 0.f;                // inf / NaN rows do not take part
  }
  val = warp_max(val);
  mw = warp_max(mw);
  if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = val; redw[threadIdx.x >> 5] = mw; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f, m = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v = fmaxf(v, redv[j]); m = fmaxf(m, redw[j]); }
    float t = v * m;
    if (!(t < 3.0e38f)) t = 3.0e38f;
    if (t > 0.f) atomicMax(reinterpret_cast<unsigned int*>(x_aux + 3 * n), __float_as_uint(t));
  }
}

// S of the staged R (see bmu_decode_stat_kernel): the power of two that puts inv_count * stat into [2^13, 2^14).
__device__ __forceinline__ float f16_r_scale(float stat, float inv_count) {
  const float t = stat * inv_count;
  if (!(t > 0.f) || !(t < 3.0e38f)) return 1.f;
  int p;
  frexpf(t, &p);                                    // t = m * 2^p, m in [0.5, 1)
  int e = F16_TOP_EXP - p;
  e = e < -120 ? -120 : (e > 120 ? 120 : e);
  return __uint_as_float(static_cast<uint32_t>(e + 127) << 23);
}

// w = exp(-|p_k - p_bmu|^2 / (2 T^2)), written the way the reference evaluates it:
// norm first, then square, divide, exp (models/som_layer.py:149-150).
__device__ __forceinline__ float neighbourhood_weight(float pky, float pkx, float pby, float pbx, float two_t2) {
  const float dy = pky - pby, dx = pkx - pbx;
  const float dist = sqrtf(dy * dy + dx * dx);
  return expf(-(dist * dist) / two_t2);
}

constexpr int COLS_PER_BLOCK = 1024;   // 256 threads x 4 columns

__global__ void __launch_bounds__(256)
neighbourhood_kernel(const long long* __restrict__ bmu, const float* __restrict__ pos, long long K, long long k_offset,
                     const float* __restrict__ T_dev, float* __restrict__ w, long long ldw) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const long long b = blockIdx.x;
  const float T = __ldg(T_dev);
  const float two_t2 = 2.f * (T * T);
  const long long bi = bmu[b];
  const float pby = __ldg(pos + 2 * bi), pbx = __ldg(pos + 2 * bi + 1);
  for (int i = 0; i < 4; ++i) {
    const long long k = static_cast<long long>(blockIdx.y) * COLS_PER_BLOCK + i * 256 + threadIdx.x;
    if (k < K) {
      const float2 pk = __ldg(reinterpret_cast<const float2*>(pos) + k_offset + k);
      w[b * ldw + k] = neighbourhood_weight(pk.x, pk.y, pby, pbx, two_t2);
    }
  }
}

__global__ void __launch_bounds__(256)
weighted_loss_kernel(const float* __restrict__ dist, long long ldd, const long long* __restrict__ bmu,
                     const float* __restrict__ pos, long long K, long long k_offset, const float* __restrict__ T_dev,
                     float inv_count, float* __restrict__ partials, float* __restrict__ loss_out) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const long long b = blockIdx.x;
  const float T = __ldg(T_dev);
  const float two_t2 = 2.f * (T * T);
  const long long bi = bmu[b];
  const float pby = __ldg(pos + 2 * bi), pbx = __ldg(pos + 2 * bi + 1);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long k = static_cast<long long>(blockIdx.y) * COLS_PER_BLOCK + i * 256 + threadIdx.x;
    if (k < K) {
      const float2 pk = __ldg(reinterpret_cast<const float2*>(pos) + k_offset + k);
      s = fmaf(neighbourhood_weight(pk.x, pk.y, pby, pbx, two_t2), __ldg(dist + b * ldd + k), s);
    }
  }
  __shared__ float red[8];
  __shared__ double dred[256];
  __shared__ bool is_last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  const long long nblocks = static_cast<long long>(gridDim.x) * gridDim.y;
  // scratch layout: word 0 = completion counter (fixed slot, independent of the grid), partial sums from word 2
  unsigned int* counter = reinterpret_cast<unsigned int*>(partials);
  partials += 2;
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    partials[static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x] = t;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == nblocks - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // Last block: fixed-order fp64 reduction of all partials -> deterministic loss.
  __threadfence();
  double acc = 0.0;
  for (long long i = threadIdx.x; i < nblocks; i += 256) acc += static_cast<double>(__ldcg(partials + i));
  dred[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *loss_out = static_cast<float>(dred[0] * static_cast<double>(inv_count));
    *counter = 0u;                                   // restore the zero state for the next call
  }
}

__global__ void __launch_bounds__(256)
loss_grad_kernel(const long long* __restrict__ bmu, const float* __restrict__ pos, long long K, long long k_offset,
                 const float* __restrict__ T_dev, const float* __restrict__ g_out, float inv_count,
                 float* __restrict__ G, long long ldg) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const long long b = blockIdx.x;
  const float T = __ldg(T_dev);
  const float two_t2 = 2.f * (T * T);
  const float scale = __ldg(g_out) * inv_count;
  const long long bi = bmu[b];
  const float pby = __ldg(pos + 2 * bi), pbx = __ldg(pos + 2 * bi + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long k = static_cast<long long>(blockIdx.y) * COLS_PER_BLOCK + i * 256 + threadIdx.x;
    if (k < K) {
      const float2 pk = __ldg(reinterpret_cast<const float2*>(pos) + k_offset + k);
      G[b * ldg + k] = scale * neighbourhood_weight(pk.x, pk.y, pby, pbx, two_t2);
    }
  }
}

constexpr int COEFF_ROWS = 32;

// thread <-> column k, loop over COEFF_ROWS rows: coalesced reads of G/dist, coalesced writes of R.
__global__ void __launch_bounds__(256)
bwd_coeffs_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ dist, long long ldd,
                  long long B, long long K, int mode, const float* __restrict__ x_aux, const float* __restrict__ w_aux,
                  float* __restrict__ r_hi, float* __restrict__ r_lo, long long ldr,
                  float* __restrict__ ax, float* __restrict__ bx, float* __restrict__ aw, float* __restrict__ bw) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const long long k = static_cast<long long>(blockIdx.y) * 256 + threadIdx.x;
  const long long b0 = static_cast<long long>(blockIdx.x) * COEFF_ROWS;
  const bool col_ok = k < K;
  __shared__ float red[COEFF_ROWS][8];
  float colsum = 0.f;
#pragma unroll 4
  for (int r = 0; r < COEFF_ROWS; ++r) {
    const long long b = b0 + r;
    float rowterm = 0.f;
    if (b < B && col_ok) {
      const float g = __ldg(G + b * ldg + k);
      const float d = __ldg(dist + b * ldd + k);
      float rv, term;
      if (mode == 0) {
        rv = (d == 0.f) ? 0.f : g / d;               // ATen: ratio.masked_fill_(dist == 0, 0)
        term = rv;
      } else {
        rv = g;
        term = g * (1.f - d);                        // g * (x̂ . ŵ): projection coefficient of normalize backward
      }
      const float h = tf32_rna(rv);
      r_hi[b * ldr + k] = h;
      r_lo[b * ldr + k] = tf32_rna(rv - h);
      colsum += term;
      rowterm = term;
    }
    rowterm = warp_sum(rowterm);
    if ((threadIdx.x & 31) == 0) red[r][threadIdx.x >> 5] = rowterm;
  }
  __syncthreads();
  if (threadIdx.x < COEFF_ROWS) {
    const long long b = b0 + threadIdx.x;
    if (b < B) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
      const float xa = (mode == 1) ? __ldg(x_aux + b) : 1.f;
      atomicAdd(ax + b, (mode == 1) ? xa * xa * t : t);
      if (blockIdx.y == 0) bx[b] = xa;
    }
  }
  if (col_ok) {
    const float wa = (mode == 1) ? __ldg(w_aux + k) : 1.f;
    atomicAdd(aw + k, (mode == 1) ? wa * wa * colsum : colsum);
    if (blockIdx.x == 0) bw[k] = wa;
  }
}

// Fused forward-loss + backward staging (models/som_layer.py:137-152 and the MeanBackward0 -> MulBackward0 ->
// distance-backward chain): one pass over dist[B,K] produces
//   * the loss  inv_count * sum w d  (w recomputed in registers, never stored),
//   * R_unit = dL/dd / d for unit upstream gradient, as the tf32 hi/lo GEMM operand, and
//   * partial row / column sums of it (the rank-1 coefficients of the closed-form backward).
// The upstream gradient g_out multiplies the result in the epilogue of the gradient GEMMs, so backward needs no
// further pass over B x K.  r_hi == nullptr: loss only (validation / no_grad).
// Block = rows_per_block rows x 512 columns: warp (rh, cs) owns half of the rows x 128 columns, a lane 4 consecutive
// columns (one 16-byte load of dist, two 16-byte stores of R per row; the loads of the next 4 rows are issued before
// the current 4 are processed).
// Everything is deterministic (run-to-run bit-identical): no atomics.  A warp writes the sum of its 128 columns of a
// row to row_part[b][slab] (slab = global 128-column slab), a block writes its column sums (the two row halves added
// in shared memory, fixed order) to col_part[row block][k]; the gradient GEMM epilogues add the partial sums of a
// row in index order (som_gemm.cuh grad_coeffs).  The loss: per-block partial, the last block to finish adds all
// partials in fp64 in a fixed order - only its first warp stays for that, nobody waits at a block barrier behind
// the completion counter.
constexpr int LC_ROWS = 8;
constexpr int LC_COLS = 512;

// SQUARE: the map is the canonical integer grid of the square topology (models/som_layer.py:61-67), cell k at
// (k / cols, k % cols).  Then the neighbourhood weight factorises, exp(-(dr^2 + dc^2) / 2T^2) = e[|dr|] * e[|dc|], and
// a block needs max(rows, cols) exponentials in shared memory instead of one per element.  The factorised weight
// differs from the reference's exp(-(sqrt(dr^2 + dc^2))^2 / 2T^2) by a few ulp (1e-7 relative), inside the 1e-5 budget.
constexpr int LC_MAX_TAB = 1024;

template <bool SQUARE>
__global__ void __launch_bounds__(256)
loss_coeffs_kernel(const float* __restrict__ dist, long long ldd, const long long* __restrict__ bmu,
                   const float* __restrict__ pos, int grid_rows, int grid_cols, long long B, long long K,
                   long long k_offset, const float* __restrict__ T_dev, float inv_count, int mode,
                   float* __restrict__ r_hi, float* __restrict__ r_lo, long long ldr,
                   float* __restrict__ row_part, int n_row_parts, float* __restrict__ col_part,
                   float* __restrict__ partials, float* __restrict__ loss_out, int rows_per_block,
                   const float* __restrict__ x_aux, const float* __restrict__ w_aux) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rh = warp >> 2, cs = warp & 3;
  // 3xFP16 (x_aux != nullptr): R is staged as fp16 hi/lo of R * 2^-e_b * 2^-g_k * S (see bmu_decode_stat_kernel); r_hi / r_lo
  // then are __half matrices with row pitch ldr halves.  1 / S goes behind the partial-sum tables for the GEMM epilogues.
  const bool f16 = x_aux != nullptr;
  float s_glob = 1.f;
  if (f16) {
    s_glob = f16_r_scale(__ldg(x_aux + 3 * B), inv_count);
    if (r_hi && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      row_part[B * n_row_parts] = 1.f / s_glob;
      col_part[static_cast<long long>(gridDim.x) * K] = 1.f / s_glob;
    }
  }
  const int half_rows = rows_per_block >> 1;              // rows handled by each of the two warp rows (multiple of 4)
  const long long k0 = static_cast<long long>(blockIdx.y) * LC_COLS + cs * 128 + lane * 4;
  const long long b_base = static_cast<long long>(blockIdx.x) * rows_per_block + rh * half_rows;
  const int slab = static_cast<int>(blockIdx.y) * 4 + cs;     // global 128-column slab of this warp
  const bool want_r = r_hi != nullptr;
  const bool vec = (ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(dist) & 15) == 0 &&
                   (!want_r || ((ldr & 3) == 0 && ((reinterpret_cast<uintptr_t>(r_hi) | reinterpret_cast<uintptr_t>(r_lo)) & 15) == 0));
  float wsc[4] = {0.f, 0.f, 0.f, 0.f};                     // fp16: 2^-g_k * S of this lane's columns
  if (f16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) if (k0 + i < K) wsc[i] = __ldg(w_aux + K + k0 + i) * s_glob;
  }
  __shared__ float lred[8];
  __shared__ float csum[LC_COLS];
  __shared__ float tab[SQUARE ? LC_MAX_TAB : 1];

  const float T = __ldg(T_dev);
  const float two_t2 = 2.f * (T * T);
  float2 pk[4];
  int rk[4], ck[4];
  bool ok[4];
  {
    // grid cell of this lane's first column (32-bit arithmetic: K < 2^31), the next three follow by increment
    int r0 = 0, c0 = 0;
    if constexpr (SQUARE) {
      const unsigned kg = static_cast<unsigned>(k_offset + k0);
      r0 = static_cast<int>(kg / static_cast<unsigned>(grid_cols));
      c0 = static_cast<int>(kg - static_cast<unsigned>(r0) * static_cast<unsigned>(grid_cols));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i] = k0 + i < K;
      pk[i] = make_float2(0.f, 0.f);
      rk[i] = r0; ck[i] = c0;
      if constexpr (SQUARE) {
        if (++c0 == grid_cols) { c0 = 0; ++r0; }
      } else if (ok[i]) {
        pk[i] = __ldg(reinterpret_cast<const float2*>(pos) + k_offset + k0 + i);
      }
    }
  }
  if constexpr (SQUARE) {
    const int ntab = max(grid_rows, grid_cols);
    for (int d = threadIdx.x; d < ntab; d += blockDim.x) {
      const float fd = static_cast<float>(d);
      tab[d] = expf(-(fd * fd) / two_t2);
    }
    __syncthreads();
  }
  // the 4 distances of row b in this lane's columns
  auto load_dist = [&](long long b, float (&d)[4]) {
    d[0] = d[1] = d[2] = d[3] = 0.f;
    if (b < B && ok[0]) {
      if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dist + b * ldd + k0));
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (ok[i]) d[i] = __ldg(dist + b * ldd + k0 + i);
      }
    }
  };
  float colsum[4] = {0.f, 0.f, 0.f, 0.f};
  float lsum = 0.f;
  float dn[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) load_dist(b_base + r, dn[r]);
  for (int r4 = 0; r4 < half_rows; r4 += 4) {
    const long long b0 = b_base + r4;
    if (b0 >= B) break;                                    // warp-uniform
    float dc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dc[r][i] = dn[r][i];
    }
    if (r4 + 4 < half_rows) {                              // next group's distances: in flight while this one is processed
#pragma unroll
      for (int r = 0; r < 4; ++r) load_dist(b0 + 4 + r, dn[r]);
    }
    // BMU grid position of these 4 rows: lane r loads row r, broadcast by shuffle below
    float2 pbv = make_float2(0.f, 0.f);
    int rbv = 0, cbv = 0;
    if (lane < 4 && b0 + lane < B) {
      const long long bi = bmu[b0 + lane];
      if constexpr (SQUARE) {
        const unsigned ub = static_cast<unsigned>(bi);
        rbv = static_cast<int>(ub / static_cast<unsigned>(grid_cols));
        cbv = static_cast<int>(ub - static_cast<unsigned>(rbv) * static_cast<unsigned>(grid_cols));
      } else {
        pbv = __ldg(reinterpret_cast<const float2*>(pos) + bi);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long b = b0 + r;
      const float pby = __shfl_sync(0xffffffffu, pbv.x, r), pbx = __shfl_sync(0xffffffffu, pbv.y, r);
      const int rb = __shfl_sync(0xffffffffu, rbv, r), cb = __shfl_sync(0xffffffffu, cbv, r);
      if (b >= B) break;                                   // warp-uniform
      const float (&d)[4] = dc[r];
      const float xsb = f16 ? __ldg(x_aux + B + b) : 1.f;
      float h[4], l[4], rowterm = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h[i] = 0.f; l[i] = 0.f;
        if (ok[i]) {
          float w;
          if constexpr (SQUARE) w = tab[abs(rk[i] - rb)] * tab[abs(ck[i] - cb)];
          else w = neighbourhood_weight(pk[i].x, pk[i].y, pby, pbx, two_t2);
          lsum = fmaf(w, d[i], lsum);
          if (want_r) {
            const float g = inv_count * w;
            float rv, term;
            if (mode == 0) {
              rv = (d[i] == 0.f) ? 0.f : __fdividef(g, d[i]);   // ATen: ratio.masked_fill_(dist == 0, 0); 2-ulp divide
              term = rv;
            } else {
              rv = g;
              term = g * (1.f - d[i]);                     // g * (x^ . w^): projection coefficient of normalize backward
            }
            if (f16) {
              h[i] = fminf((rv * xsb) * wsc[i], 60000.f);   // scaled value; split below
            } else {
              h[i] = tf32_rna(rv);
              l[i] = tf32_rna(rv - h[i]);
            }
            colsum[i] += term;
            rowterm += term;
          }
        }
      }
      if (want_r && f16) {
        __half* hp16 = reinterpret_cast<__half*>(r_hi) + b * ldr + k0;
        __half* lp16 = reinterpret_cast<__half*>(r_lo) + b * ldr + k0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ok[i]) {
            __half hh, ll;
            split_f16(h[i], hh, ll);
            hp16[i] = hh; lp16[i] = ll;
          }
        }
        rowterm = warp_sum(rowterm);
        if (lane == 0 && slab < n_row_parts) row_part[b * n_row_parts + slab] = rowterm;
      } else if (want_r) {
        if (ok[0]) {
          if (vec) {
            *reinterpret_cast<float4*>(r_hi + b * ldr + k0) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(r_lo + b * ldr + k0) = make_float4(l[0], l[1], l[2], l[3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (ok[i]) { r_hi[b * ldr + k0 + i] = h[i]; r_lo[b * ldr + k0 + i] = l[i]; }
          }
        }
        rowterm = warp_sum(rowterm);                       // xor butterfly: the same value, bit for bit, in every run
        if (lane == 0 && slab < n_row_parts) row_part[b * n_row_parts + slab] = rowterm;
      }
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0) lred[warp] = lsum;
  if (want_r && rh == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) csum[cs * 128 + lane * 4 + i] = colsum[i];
  }
  __syncthreads();
  if (want_r && rh == 0) {
    float* cp = col_part + static_cast<long long>(blockIdx.x) * K + k0;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (ok[i]) cp[i] = colsum[i] + csum[cs * 128 + lane * 4 + i];
  }
  if (warp != 0) return;
  // deterministic loss: per-block partial, the last block's first warp adds all partials in a fixed order in fp64
  const long long nblocks = static_cast<long long>(gridDim.x) * gridDim.y;
  unsigned int* counter = reinterpret_cast<unsigned int*>(partials);
  partials += 2;
  unsigned int done = 0;
  if (lane == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += lred[i];
    partials[static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x] = t;
    __threadfence();
    done = atomicAdd(counter, 1u);
  }
  done = __shfl_sync(0xffffffffu, done, 0);
  if (done != nblocks - 1) return;
  __threadfence();
  double acc = 0.0;
  for (long long i = lane; i < nblocks; i += 32) acc += static_cast<double>(__ldcg(partials + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    *loss_out = static_cast<float>(acc * static_cast<double>(inv_count));
    *counter = 0u;                                   // restore the zero state for the next call
  }
}

// Fast path of the fused loss kernel for what the training step of a square map always presents: 16-byte aligned
// rows, K and the grid's column count multiples of 4 (a lane's 4 columns then lie in ONE grid row), backward staging
// requested.  Same block / warp / lane decomposition and the same outputs as loss_coeffs_kernel<true>, but straight-line
// code: the generic kernel spends ~70 instructions per element (ncu: 2.7 warp instructions per cycle and SM, issue
// bound at 3.9 TB/s on a 4096 x 16384 chunk) on per-element predicates, 64-bit index arithmetic and two table lookups
// per element; here the row factor of the weight is looked up once per lane and row, inv_count is folded into it,
// pointers advance by increments and nothing is predicated per element.
template <int MODE, bool F16>
__global__ void __launch_bounds__(256)
loss_coeffs_fast_kernel(const float* __restrict__ dist, long long ldd, const long long* __restrict__ bmu,
                        int grid_rows, int grid_cols, long long B, long long K, long long k_offset,
                        const float* __restrict__ T_dev, float inv_count,
                        float* __restrict__ r_hi, float* __restrict__ r_lo, long long ldr,
                        float* __restrict__ row_part, int n_row_parts, float* __restrict__ col_part,
                        float* __restrict__ partials, float* __restrict__ loss_out, int rows_per_block,
                        const float* __restrict__ x_aux, const float* __restrict__ w_aux) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  // F16: R staged as fp16 hi/lo of R * 2^-e_b * 2^-g_k * S (see loss_coeffs_kernel / bmu_decode_stat_kernel)
  float s_glob = 1.f;
  if constexpr (F16) {
    s_glob = f16_r_scale(__ldg(x_aux + 3 * B), inv_count);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      row_part[B * n_row_parts] = 1.f / s_glob;
      col_part[static_cast<long long>(gridDim.x) * K] = 1.f / s_glob;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rh = warp >> 2, cs = warp & 3;
  const int half_rows = rows_per_block >> 1;
  const long long k0 = static_cast<long long>(blockIdx.y) * LC_COLS + cs * 128 + lane * 4;
  const long long b_base = static_cast<long long>(blockIdx.x) * rows_per_block + rh * half_rows;
  const int slab = static_cast<int>(blockIdx.y) * 4 + cs;
  const bool col_ok = k0 < K;                              // K % 4 == 0: all four columns of the lane or none
  __shared__ float lred[8];
  __shared__ float csum[LC_COLS];
  __shared__ float tab[LC_MAX_TAB];
  __shared__ float tab_s[LC_MAX_TAB];                      // inv_count * tab: the row factor carries the 1 / (B K)

  const float T = __ldg(T_dev);
  const float two_t2 = 2.f * (T * T);
  const unsigned kg = static_cast<unsigned>(k_offset + k0);
  const int rk = static_cast<int>(kg / static_cast<unsigned>(grid_cols));
  const int ck = static_cast<int>(kg - static_cast<unsigned>(rk) * static_cast<unsigned>(grid_cols));
  {
    const int ntab = max(grid_rows, grid_cols);
    for (int d = threadIdx.x; d < ntab; d += blockDim.x) {
      const float fd = static_cast<float>(d);
      const float e = expf(-(fd * fd) / two_t2);
      tab[d] = e;
      tab_s[d] = inv_count * e;
    }
    __syncthreads();
  }
  const long long rows_left = B - b_base;                  // rows of this warp that exist (may be <= 0)
  const int nrows = rows_left <= 0 ? 0 : (rows_left < half_rows ? static_cast<int>(rows_left) : half_rows);
  const float* dp = dist + b_base * ldd + k0;
  float* hp = r_hi + b_base * ldr + k0;
  float* lp = r_lo + b_base * ldr + k0;
  __half* hp16 = reinterpret_cast<__half*>(r_hi) + b_base * ldr + k0;      // F16: half matrices, pitch ldr halves
  __half* lp16 = reinterpret_cast<__half*>(r_lo) + b_base * ldr + k0;
  float4 csc = make_float4(0.f, 0.f, 0.f, 0.f);                            // F16: 2^-g_k * S of this lane's 4 columns
  if constexpr (F16) {
    if (col_ok) {
      csc = __ldg(reinterpret_cast<const float4*>(w_aux + K + k0));
      csc.x *= s_glob; csc.y *= s_glob; csc.z *= s_glob; csc.w *= s_glob;
    }
  }
  float* rp = row_part + b_base * n_row_parts + slab;
  const bool slab_ok = slab < n_row_parts;
  auto load4 = [&](int r, float4& v) {
    v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && r < nrows) v = __ldg(reinterpret_cast<const float4*>(dp + static_cast<long long>(r) * ldd));   // r: row of the warp
  };
  float colsum0 = 0.f, colsum1 = 0.f, colsum2 = 0.f, colsum3 = 0.f, lsum = 0.f;
  // One group = 4 rows.  All arithmetic runs unconditionally (missing rows / columns contribute exact zeros: their
  // distances load as 0 and their row factor is 0), only the stores are predicated.
  auto process = [&](const float4 (&dg)[4], int r4) {
    // BMU cell of these 4 rows: lane r loads row r, one shuffle per row broadcasts (row << 16 | column)
    int cell = 0;
    float xsl = 0.f;                                       // F16: 2^-e of the row this lane fetched
    if (lane < 4 && r4 + lane < nrows) {
      const unsigned ub = static_cast<unsigned>(bmu[b_base + r4 + lane]);
      const unsigned rb = ub / static_cast<unsigned>(grid_cols);
      cell = static_cast<int>((rb << 16) | (ub - rb * static_cast<unsigned>(grid_cols)));
      if constexpr (F16) xsl = __ldg(x_aux + B + b_base + r4 + lane);
    }
    float t[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rc = __shfl_sync(0xffffffffu, cell, r);
      const bool row_ok = r4 + r < nrows;                  // warp-uniform
      const int rb = rc >> 16, cb = rc & 0xffff;
      const float er = row_ok ? tab_s[abs(rk - rb)] : 0.f;  // inv_count * exp(-dr^2 / 2T^2)
      const int dc0 = ck - cb;
      const float w0 = er * tab[abs(dc0)], w1 = er * tab[abs(dc0 + 1)], w2 = er * tab[abs(dc0 + 2)], w3 = er * tab[abs(dc0 + 3)];
      const float4 d = dg[r];
      lsum = fmaf(w0, d.x, lsum); lsum = fmaf(w1, d.y, lsum); lsum = fmaf(w2, d.z, lsum); lsum = fmaf(w3, d.w, lsum);
      float4 rv, term;
      if (MODE == 0) {
        // ATen: ratio = grad / dist, masked_fill_(dist == 0, 0).  w * rcp(d): 2 ulp; rcp(0) = inf is selected away
        rv.x = d.x == 0.f ? 0.f : w0 * fast_rcp(d.x);
        rv.y = d.y == 0.f ? 0.f : w1 * fast_rcp(d.y);
        rv.z = d.z == 0.f ? 0.f : w2 * fast_rcp(d.z);
        rv.w = d.w == 0.f ? 0.f : w3 * fast_rcp(d.w);
        term = rv;
      } else {
        rv = make_float4(w0, w1, w2, w3);
        term = make_float4(w0 * (1.f - d.x), w1 * (1.f - d.y), w2 * (1.f - d.z), w3 * (1.f - d.w));
      }
      if constexpr (F16) {
        const float xsb = __shfl_sync(0xffffffffu, xsl, r);
        float4 q;
        q.x = fminf((rv.x * xsb) * csc.x, 60000.f); q.y = fminf((rv.y * xsb) * csc.y, 60000.f);
        q.z = fminf((rv.z * xsb) * csc.z, 60000.f); q.w = fminf((rv.w * xsb) * csc.w, 60000.f);
        if (col_ok && row_ok) store_split4_f16(hp16, lp16, 0, q, 1.f);
        hp16 += ldr; lp16 += ldr;
      } else {
        float4 h, l;
        h.x = tf32_rna_finite(rv.x); h.y = tf32_rna_finite(rv.y); h.z = tf32_rna_finite(rv.z); h.w = tf32_rna_finite(rv.w);
        l.x = tf32_rna_finite(rv.x - h.x); l.y = tf32_rna_finite(rv.y - h.y);
        l.z = tf32_rna_finite(rv.z - h.z); l.w = tf32_rna_finite(rv.w - h.w);
        if (col_ok && row_ok) {
          *reinterpret_cast<float4*>(hp) = h;
          *reinterpret_cast<float4*>(lp) = l;
        }
        hp += ldr; lp += ldr;
      }
      colsum0 += term.x; colsum1 += term.y; colsum2 += term.z; colsum3 += term.w;
      t[r] = col_ok ? (term.x + term.y) + (term.z + term.w) : 0.f;
    }
    // the four row sums of the group in one transposed butterfly (6 shuffles instead of 20; a fixed pattern, so the
    // result is the same in every run): after the 16- and 8-steps a lane carries ONE of the rows, selected by its bits
    // 4 and 3; lanes 0, 16, 8, 24 end up with the totals of rows 0, 1, 2, 3
    const bool b16 = (lane & 16) != 0, b8 = (lane & 8) != 0;
    const float v01 = (b16 ? t[1] : t[0]) + __shfl_xor_sync(0xffffffffu, b16 ? t[0] : t[1], 16);
    const float v23 = (b16 ? t[3] : t[2]) + __shfl_xor_sync(0xffffffffu, b16 ? t[2] : t[3], 16);
    float u = (b8 ? v23 : v01) + __shfl_xor_sync(0xffffffffu, b8 ? v01 : v23, 8);
    u += __shfl_xor_sync(0xffffffffu, u, 4);
    u += __shfl_xor_sync(0xffffffffu, u, 2);
    u += __shfl_xor_sync(0xffffffffu, u, 1);
    const int row = (b16 ? 1 : 0) + (b8 ? 2 : 0);
    if ((lane & 7) == 0 && slab_ok && r4 + row < nrows) rp[static_cast<long long>(row) * n_row_parts] = u;
    rp += 4 * n_row_parts;
  };
  // two register sets of distances alternate (no copies): while one group is processed the next one is in flight
  float4 da[4], db[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) load4(r, da[r]);
  for (int r4 = 0; r4 < nrows; r4 += 8) {
#pragma unroll
    for (int r = 0; r < 4; ++r) load4(r4 + 4 + r, db[r]);
    process(da, r4);
    if (r4 + 4 >= nrows) break;
#pragma unroll
    for (int r = 0; r < 4; ++r) load4(r4 + 8 + r, da[r]);
    process(db, r4 + 4);
  }
  if (!col_ok) { colsum0 = colsum1 = colsum2 = colsum3 = 0.f; lsum = 0.f; }
  lsum = warp_sum(lsum);
  if (lane == 0) lred[warp] = lsum;
  if (rh == 1) {
    float* c = csum + cs * 128 + lane * 4;
    c[0] = colsum0; c[1] = colsum1; c[2] = colsum2; c[3] = colsum3;
  }
  __syncthreads();
  if (rh == 0 && col_ok) {
    const float* c = csum + cs * 128 + lane * 4;
    *reinterpret_cast<float4*>(col_part + static_cast<long long>(blockIdx.x) * K + k0) =
        make_float4(colsum0 + c[0], colsum1 + c[1], colsum2 + c[2], colsum3 + c[3]);
  }
  if (warp != 0) return;
  // deterministic loss: per-block partial (already scaled by inv_count), the last block's first warp adds all partials
  // in a fixed order in fp64
  const long long nblocks = static_cast<long long>(gridDim.x) * gridDim.y;
  unsigned int* counter = reinterpret_cast<unsigned int*>(partials);
  partials += 2;
  unsigned int done = 0;
  if (lane == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += lred[i];
    partials[static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x] = t;
    __threadfence();
    done = atomicAdd(counter, 1u);
  }
  done = __shfl_sync(0xffffffffu, done, 0);
  if (done != nblocks - 1) return;
  __threadfence();
  double acc = 0.0;
  for (long long i = lane; i < nblocks; i += 32) acc += static_cast<double>(__ldcg(partials + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    *loss_out = static_cast<float>(acc);
    *counter = 0u;
  }
}

// Rows per block of the fused loss kernel: as many as keep ~8 blocks per SM in flight (fewer, taller blocks amortise
// the per-block setup and shorten the column-sum tables), between LC_ROWS and 128.
inline int loss_rows_per_block(int64_t B, int64_t K, int sms) {
  const int64_t col_blocks = (K + LC_COLS - 1) / LC_COLS;
  int rpb = LC_ROWS;
  while (rpb < 128 && ((B + 2 * rpb - 1) / (2 * rpb)) * col_blocks >= 8ll * sms) rpb *= 2;
  return rpb;
}

// ----------------------------------------------------------------------------------------------
// Fused prototype optimizer step (the reference optimises som_layer.parameters() with torch.optim.AdamW,
// models/vit_som.py:140-151: default weight decay 0.01, lr and betas from the YAML) + operand staging for the next
// forward: ONE pass over W, dW, m, v that writes W, m, v and the tf32 hi/lo split (+ row norms / reciprocal norms) of
// the NEW prototypes, so that the next step's staging kernel handles the latents only.
// torch.optim.AdamW semantics (non-amsgrad), per element:
//   w *= 1 - lr * wd;  m = lerp(m, g, 1 - b1);  v = b2 * v + (1 - b2) g^2;
//   w -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),  bc1 = 1 - b1^t, bc2 = 1 - b2^t
// The step count t and the learning rate live in device memory (hp_dev), so a captured CUDA graph replays correctly:
//   hp_dev[0] = lr, hp_dev[1] = t (the 1-based step count as a float; the caller advances it on the device before the
//   call), hp_dev[2] = gradient scale (e.g. 1 / world for a summed gradient; 1 otherwise).
// Block per row (256 threads), rows of up to 8192 elements stay in registers between the update and the staging.
// ----------------------------------------------------------------------------------------------
struct AdamHyper { double beta1, beta2, eps, weight_decay; };   // doubles, like torch's Python scalars

// GROUP threads cooperate on one row (GROUP = 32: warp per row for short rows, GROUP = 256: block per row).
template <int GROUP, bool F16>
__global__ void __launch_bounds__(256, 2)
adamw_stage_kernel(float* __restrict__ W, long long ldw, const float* __restrict__ dW, long long lddw,
                   float* __restrict__ m, float* __restrict__ v, long long ldm, long long rows, int dim,
                   const float* __restrict__ hp_dev, AdamHyper hp, int mode,
                   float* __restrict__ hi, float* __restrict__ lo, long long ld_out, float* __restrict__ aux) {
  som::pdl_wait();
  som::pdl_launch_dependents();
  constexpr int ROWS_PER_BLOCK = 256 / GROUP;
  const int gi = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
  const long long row = static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK + gi;
  const bool active = row < rows;
  const float lr = __ldg(hp_dev), t = __ldg(hp_dev + 1), gscale = __ldg(hp_dev + 2);
  // bias corrections in double like torch's Python scalars (bias_correction2_sqrt = sqrt(1 - beta2^t))
  const double bc1 = 1.0 - pow(hp.beta1, static_cast<double>(t));
  const double bc2 = 1.0 - pow(hp.beta2, static_cast<double>(t));
  const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  const float bc2_sqrt = static_cast<float>(sqrt(bc2));
  const float decay = static_cast<float>(1.0 - static_cast<double>(lr) * hp.weight_decay);
  const float omb1 = static_cast<float>(1.0 - hp.beta1), omb2 = static_cast<float>(1.0 - hp.beta2);
  const float b2 = static_cast<float>(hp.beta2), eps = static_cast<float>(hp.eps);
  const long long r_ = active ? row : 0;
  float* wrow = W + r_ * ldw;
  const float* grow = dW + r_ * lddw;
  float* mrow = m + r_ * ldm;
  float* vrow = v + r_ * ldm;
  auto update = [&](float w, float g, float& mm, float& vv) {
    g *= gscale;
    w *= decay;
    mm = fmaf(omb1, g - mm, mm);                             // lerp(m, g, 1 - b1)
    vv = fmaf(omb2, g * g, b2 * vv);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    return w - step_size * (mm / denom);
  };
  const bool vec = (dim & 3) == 0 && (ldw & 3) == 0 && (lddw & 3) == 0 && (ldm & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(dW) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  constexpr int CACHE = 8;
  const bool cached = vec && dim <= CACHE * GROUP * 4;      // the new row waits in registers for its norm
  float4 cache[CACHE];
  float ss = 0.f, amax = 0.f;                                // amax: largest magnitude of the new row (fp16 staging)
  if (active) {
    if (cached) {
      // four row segments at a time: all their loads first (4 arrays x 4 x 16 bytes in flight per thread), then the
      // arithmetic and the stores
#pragma unroll
      for (int j0 = 0; j0 < CACHE; j0 += 4) {
        float4 w4[4], g4[4], m4[4], v4[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = (gt + (j0 + jj) * GROUP) * 4;
          if (i < dim) {
            w4[jj] = *reinterpret_cast<const float4*>(wrow + i);
            g4[jj] = __ldg(reinterpret_cast<const float4*>(grow + i));
            m4[jj] = *reinterpret_cast<const float4*>(mrow + i);
            v4[jj] = *reinterpret_cast<const float4*>(vrow + i);
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = (gt + (j0 + jj) * GROUP) * 4;
          cache[j0 + jj] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < dim) {
            float4 w = w4[jj], mm = m4[jj], vv = v4[jj];
            const float4 g = g4[jj];
            w.x = update(w.x, g.x, mm.x, vv.x); w.y = update(w.y, g.y, mm.y, vv.y);
            w.z = update(w.z, g.z, mm.z, vv.z); w.w = update(w.w, g.w, mm.w, vv.w);
            *reinterpret_cast<float4*>(wrow + i) = w;
            *reinterpret_cast<float4*>(mrow + i) = mm;
            *reinterpret_cast<float4*>(vrow + i) = vv;
            ss = fmaf(w.x, w.x, ss); ss = fmaf(w.y, w.y, ss); ss = fmaf(w.z, w.z, ss); ss = fmaf(w.w, w.w, ss);
            if constexpr (F16) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(w.x), fabsf(w.y))), fmaxf(fabsf(w.z), fabsf(w.w)));
            cache[j0 + jj] = w;
          }
        }
      }
    } else if (vec) {
      for (int i = gt * 4; i < dim; i += GROUP * 4) {
        float4 w = *reinterpret_cast<const float4*>(wrow + i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(grow + i));
        float4 mm = *reinterpret_cast<const float4*>(mrow + i);
        float4 vv = *reinterpret_cast<const float4*>(vrow + i);
        w.x = update(w.x, g.x, mm.x, vv.x); w.y = update(w.y, g.y, mm.y, vv.y);
        w.z = update(w.z, g.z, mm.z, vv.z); w.w = update(w.w, g.w, mm.w, vv.w);
        *reinterpret_cast<float4*>(wrow + i) = w;
        *reinterpret_cast<float4*>(mrow + i) = mm;
        *reinterpret_cast<float4*>(vrow + i) = vv;
        ss = fmaf(w.x, w.x, ss); ss = fmaf(w.y, w.y, ss); ss = fmaf(w.z, w.z, ss); ss = fmaf(w.w, w.w, ss);
        if constexpr (F16) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(w.x), fabsf(w.y))), fmaxf(fabsf(w.z), fabsf(w.w)));
      }
    } else {
      for (int i = gt; i < dim; i += GROUP) {
        float mm = mrow[i], vv = vrow[i];
        const float w = update(wrow[i], __ldg(grow + i), mm, vv);
        wrow[i] = w; mrow[i] = mm; vrow[i] = vv;
        ss = fmaf(w, w, ss);
        if constexpr (F16) amax = fmaxf(amax, fabsf(w));
      }
    }
  }
  if (!hi) return;                                           // optimizer step only (no staging requested): uniform
  ss = warp_sum(ss);
  if constexpr (F16) amax = warp_max(amax);
  if constexpr (GROUP == 256) {
    __shared__ float red[8];
    __shared__ float redm[8];
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = ss; redm[threadIdx.x >> 5] = amax; }
    __syncthreads();
    ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += red[i];
    if constexpr (F16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) amax = fmaxf(amax, redm[i]);
    }
  }
  if (!active) return;
  float denom = 1.f;
  if (mode == 1) {
    denom = fmaxf(sqrtf(ss), 1e-12f);                        // F.normalize(p=2, eps=1e-12), as prep_rows_kernel
    if (gt == 0) aux[row] = 1.f / denom;
  } else if (gt == 0) {
    aux[row] = ss;
  }
  if constexpr (F16) {
    // the row-scaled fp16 split of the NEW prototypes, exactly what prep_rows_kernel<GROUP, true> would stage
    const float scale = f16_row_scale(mode == 1 ? amax / denom : amax);
    if (gt == 0) { aux[rows + row] = 1.f / scale; aux[2 * rows + row] = scale; }
    __half* ph16 = reinterpret_cast<__half*>(hi) + row * ld_out;
    __half* pl16 = reinterpret_cast<__half*>(lo) + row * ld_out;
    const int dim_out16 = static_cast<int>(ld_out);
    auto norm4 = [&](float4 w) {
      if (mode == 1) { w.x = w.x / denom; w.y = w.y / denom; w.z = w.z / denom; w.w = w.w / denom; }
      return w;
    };
    if (cached) {
#pragma unroll
      for (int j = 0; j < CACHE; ++j) {
        const int i = (gt + j * GROUP) * 4;
        if (i < dim_out16) store_split4_f16(ph16, pl16, i, i < dim ? norm4(cache[j]) : make_float4(0.f, 0.f, 0.f, 0.f), scale);
      }
    } else if (vec) {
      for (int i = gt * 4; i < dim_out16; i += GROUP * 4)
        store_split4_f16(ph16, pl16, i, i < dim ? norm4(*reinterpret_cast<const float4*>(wrow + i)) : make_float4(0.f, 0.f, 0.f, 0.f), scale);
    } else {
      for (int i = gt; i < dim_out16; i += GROUP) {
        __half h = __float2half_rn(0.f), l = h;
        if (i < dim) {
          float w = wrow[i];
          if (mode == 1) w = w / denom;
          split_f16(w * scale, h, l);
        }
        ph16[i] = h;
        pl16[i] = l;
      }
    }
    return;
  }
  float* ph = hi + row * ld_out;
  float* pl = lo + row * ld_out;
  const int dim_out = static_cast<int>(ld_out);
  auto split1 = [&](float x, float& h, float& l) {
    if (mode == 1) x = x / denom;
    h = tf32_rna(x);
    l = tf32_rna(x - h);
  };
  auto split4 = [&](float4 w, float4& h, float4& l) {
    split1(w.x, h.x, l.x); split1(w.y, h.y, l.y); split1(w.z, h.z, l.z); split1(w.w, h.w, l.w);
  };
  if (cached) {
#pragma unroll
    for (int j = 0; j < CACHE; ++j) {
      const int i = (gt + j * GROUP) * 4;
      if (i < dim_out) {
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
        if (i < dim) split4(cache[j], h, l);
        *reinterpret_cast<float4*>(ph + i) = h;
        *reinterpret_cast<float4*>(pl + i) = l;
      }
    }
  } else if (vec) {
    for (int i = gt * 4; i < dim_out; i += GROUP * 4) {
      float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
      if (i < dim) split4(*reinterpret_cast<const float4*>(wrow + i), h, l);   // this thread's own store of the update pass
      *reinterpret_cast<float4*>(ph + i) = h;
      *reinterpret_cast<float4*>(pl + i) = l;
    }
  } else {
    for (int i = gt; i < dim_out; i += GROUP) {
      float h = 0.f, l = 0.f;
      if (i < dim) split1(wrow[i], h, l);
      ph[i] = h;
      pl[i] = l;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Prototype-gradient all-reduce over NVLink SHARP (NVLS): batch-sharded data parallelism averages dW[K, D] over the
// ranks (DDP semantics, train_vit_som.py:45,86-87 via Lightning).  Every rank's dW lives at the same offset of a
// symmetric allocation that the NVSwitch exposes as ONE multicast address.  Two-shot: rank r owns slice r of the
// gradient; multimem.ld_reduce pulls that slice from all replicas, summed inside the switch; the mean goes back to
// all replicas with one multimem.st.  Per rank 2 * n / world floats cross its links (NCCL's ring moves ~2 n), and
// there is one kernel, one barrier before and one after, instead of a multi-step protocol.
// Barrier: flags[world][channels][world] words in a second symmetric allocation; block b of rank r raises
// flag[b][r] in every peer's copy (CAS 0 -> 1, release.sys) and lowers its own copy's flag[b][peer] (CAS 1 -> 0,
// acquire.sys): self-resetting, so launches (and CUDA-graph replays) can follow each other without host involvement.
// ----------------------------------------------------------------------------------------------
// Two blocks per SM: beside a GEMM the grid is twice the number of SMs the GEMM leaves free (with more blocks than free
// SMs the late ones would only start after the GEMM, and the exchange would finish after it instead of under it).
// Measured (2 GPUs, 20 MB): 16 blocks keep ~1 MB in flight per GPU and need 75 us - latency bound, NVLink is idle.
constexpr int NVLS_MAX_BLOCKS = 64;      // flag buffers are sized for this many barrier channels; the caller picks the grid:
                                         // two blocks per SM the GEMM beside it leaves free (24 for 136 of 148 SMs), more
                                         // when nothing runs concurrently (the kernel is bound by the bytes it keeps in flight)
constexpr int NVLS_UNROLL = 8;
constexpr int NVLS_THREADS = 512;

__device__ __forceinline__ void nvls_signal(unsigned int* addr) {
  const long long t0 = clock64();
  unsigned int seen;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(seen) : "l"(addr) : "memory");
    if (seen != 0u && clock64() - t0 > 6000000000LL) { printf("som_b200: NVLS barrier signal timed out\n"); __trap(); }
  } while (seen != 0u);
}
__device__ __forceinline__ void nvls_wait(unsigned int* addr) {
  const long long t0 = clock64();
  unsigned int seen;
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(seen) : "l"(addr) : "memory");
    if (seen != 1u && clock64() - t0 > 6000000000LL) { printf("som_b200: NVLS barrier wait timed out\n"); __trap(); }
  } while (seen != 1u);
}
// All blocks with the same blockIdx on all ranks meet.  flag_ptrs[p] = rank p's flag buffer (peer mapped).
__device__ __forceinline__ void nvls_block_barrier(unsigned int* const* flag_ptrs, int rank, int world) {
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    nvls_signal(flag_ptrs[peer] + (blockIdx.x * world + rank));
    nvls_wait(flag_ptrs[rank] + (blockIdx.x * world + peer));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NVLS_THREADS, 2)
nvls_allreduce_mean_kernel(float* mc, unsigned int* const* flag_ptrs, long long n4, int rank, int world, float scale) {
  nvls_block_barrier(flag_ptrs, rank, world);            // every rank's dW is complete and visible
  const long long per = (n4 + world - 1) / world;        // float4 elements per rank slice
  const long long begin = per * rank, end = begin + per < n4 ? begin + per : n4;
  // NVLS_UNROLL independent 16-byte reductions in flight per thread: a multimem.ld_reduce is a round trip through
  // the switch (~2-3 us), so the bytes in flight set the throughput (one per thread gave ~65 GB/s)
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = begin + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < end;
       i0 += stride * NVLS_UNROLL) {
    float4 v[NVLS_UNROLL];
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < end)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < end)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(mc + 4 * i), "f"(v[u].x * scale), "f"(v[u].y * scale), "f"(v[u].z * scale), "f"(v[u].w * scale)
                     : "memory");
    }
  }
  __threadfence_system();
  nvls_block_barrier(flag_ptrs, rank, world);            // every slice has reached every replica
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// mode: distance mode (bit 0) | SOM_PREC_FP16X3
int launch_prep(const PrepSet& a, const PrepSet& b, int64_t dim, int mode, int64_t ld_out, cudaStream_t st) {
  const int per_block = dim <= 1024 ? 8 : 1;
  const long long blocks_a = (a.rows + per_block - 1) / per_block;
  const long long blocks_b = b.src ? (b.rows + per_block - 1) / per_block : 0;
  if (blocks_a + blocks_b > 0x7fffffffLL) return fail(SOM_ERR_ARG, "too many rows to stage in one launch");
  const unsigned grid = static_cast<unsigned>(blocks_a + blocks_b);
  const int dmode = mode & 1;
  const long long ld = ld_out;
  if (mode & SOM_PREC_FP16X3) {
    if (dim <= 1024) SOM_CUDA(launch_kernel(prep_rows_kernel<32, true>, dim3(grid), dim3(256), 0, st, a, b, blocks_a, static_cast<int>(dim), dmode, ld));
    else             SOM_CUDA(launch_kernel(prep_rows_kernel<256, true>, dim3(grid), dim3(256), 0, st, a, b, blocks_a, static_cast<int>(dim), dmode, ld));
  } else {
    if (dim <= 1024) SOM_CUDA(launch_kernel(prep_rows_kernel<32, false>, dim3(grid), dim3(256), 0, st, a, b, blocks_a, static_cast<int>(dim), dmode, ld));
    else             SOM_CUDA(launch_kernel(prep_rows_kernel<256, false>, dim3(grid), dim3(256), 0, st, a, b, blocks_a, static_cast<int>(dim), dmode, ld));
  }
  g_launches.fetch_add(1);
  return SOM_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int som_b200_abi_version(void) { return 7; }
const char* som_last_error(void) { return g_last_error.c_str(); }
int64_t som_launch_count(void) { return g_launches.load(); }
void som_launch_count_reset(void) { g_launches.store(0); }
void som_set_tuning(int bn_override, int kchunk) {
  g_bn_override.store(bn_override);
  if (kchunk > 0) g_kchunk.store(kchunk);
}
void som_set_debug(int bits) {
  g_debug.store(bits & ~(16 | 32));
  g_tma3d.store((bits & 16) ? 0 : 1);
  g_loss_fast.store((bits & 32) ? 0 : 1);
}
void som_set_debug_times(unsigned long long* dev_buf) { g_dbg_times.store(dev_buf); }
void som_set_pdl(int on) { g_pdl.store(on ? 1 : 0); }
void som_set_streamk(int mode) { g_streamk.store(mode < 0 ? -1 : (mode > 0 ? 1 : 0)); }
int64_t som_gemm_workspace_floats(void) {
  // 256 x 256 partial tiles for every CTA pair of the current device (148 SMs -> 74 pairs on a B200)
  DeviceInfo di;
  const int64_t pairs = device_info(di) == SOM_OK && di.sms >= 2 ? di.sms / 2 : 74;
  // (two slots per pair: the two-phase schedule of the data-parallel backward hands over one partial tile per phase)
  return SK_FLAG_WORDS + 2 * std::min<int64_t>(pairs, SK_FLAG_WORDS / 32) * 256 * 256;
}
void som_set_cta_group(int cg) { g_cg_override.store(cg == 1 || cg == 2 ? cg : 0); }

int som_prep_rows(const float* src, int64_t rows, int64_t dim, int64_t ld_src, int mode, float* hi, float* lo,
                  int64_t ld_out, float* aux, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!src || !hi || !lo || !aux) return fail(SOM_ERR_ARG, "som_prep_rows: null pointer");
  if (rows <= 0 || dim <= 0 || dim > (1ll << 30) || ld_src < dim || ld_out < dim)
    return fail(SOM_ERR_ARG, "som_prep_rows: bad shape");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_prep_rows: bad mode");
  if ((ld_out & (mode_f16(mode) ? 7 : 3)) != 0 || ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) != 0)
    return fail(SOM_ERR_ARG, "som_prep_rows: hi/lo must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  PrepSet a{src, rows, ld_src, hi, lo, aux, nullptr}, none{};
  return launch_prep(a, none, dim, mode, ld_out, as_stream(stream));
}

int som_bmu_init(long long* packed, int64_t B, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!packed || B <= 0) return fail(SOM_ERR_ARG, "som_bmu_init: bad argument");
  bmu_init_kernel<<<static_cast<unsigned>((B + 255) / 256), 256, 0, as_stream(stream)>>>(packed, B);
  SOM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SOM_OK;
}

int som_fwd_distances(const float* x_hi, const float* x_lo, int64_t ldx, const float* x_aux, const float* w_hi,
                      const float* w_lo, int64_t ldw, const float* w_aux, int64_t B, int64_t K, int64_t D, int mode,
                      int64_t idx_offset, float* dist, int64_t ldd, long long* packed, float* ws, int64_t ws_floats,
                      void* stream) {
  if (!packed) return fail(SOM_ERR_ARG, "som_fwd_distances: packed must not be null");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_fwd_distances: bad mode");
  const int f16 = mode_f16(mode), dmode = mode & 1;
  if ((dmode == SOM_MODE_EUCLIDEAN || f16) && (!x_aux || !w_aux)) return fail(SOM_ERR_ARG, "som_fwd_distances: aux vectors required");
  if (dist && ldd < K) return fail(SOM_ERR_ARG, "som_fwd_distances: ldd < K");
  if (idx_offset < 0 || idx_offset + K > 0x7fffffffLL) return fail(SOM_ERR_ARG, "som_fwd_distances: index range");
  som::EpiParams e{};
  e.row_aux = x_aux; e.col_aux = w_aux; e.dist = dist; e.ldd = ldd; e.packed = packed;
  e.idx_offset = static_cast<int>(idx_offset); e.mode = dmode;
  if (f16) { e.row_scale = x_aux + B; e.col_scale = w_aux + K; }        // the 2^-e of the staged rows
  return launch_gemm(som::EPI_DIST, x_hi, x_lo, ldx, 0, w_hi, w_lo, ldw, 0, B, K, D, 0, 0, 3, e, ws, ws_floats,
                     as_stream(stream), 0, f16);
}

int som_bmu_decode(const long long* packed, int64_t B, int64_t K_total, int64_t* bmu, float* min_key, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!packed || !bmu || B <= 0 || K_total <= 0) return fail(SOM_ERR_ARG, "som_bmu_decode: bad argument");
  SOM_CUDA(launch_kernel(bmu_decode_kernel, dim3(static_cast<unsigned>((B + 255) / 256)), dim3(256), 0, as_stream(stream),
                         packed, static_cast<long long>(B), static_cast<long long>(K_total),
                         reinterpret_cast<long long*>(bmu), min_key));
  g_launches.fetch_add(1);
  return SOM_OK;
}

int som_bmu_decode_scaled(const long long* packed, int64_t B, int64_t K_total, int64_t* bmu, float* min_key,
                          float* x_aux, const float* w_aux, int64_t K, int mode, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!packed || B <= 0 || K_total <= 0 || K <= 0 || !x_aux || !w_aux || !mode_ok(mode) || !mode_f16(mode))
    return fail(SOM_ERR_ARG, "som_bmu_decode_scaled: bad argument (fp16 stagings of this forward required)");
  SOM_CUDA(launch_kernel(bmu_decode_stat_kernel, dim3(static_cast<unsigned>((B + 255) / 256)), dim3(256), 0, as_stream(stream),
                         packed, static_cast<long long>(B), static_cast<long long>(K_total),
                         reinterpret_cast<long long*>(bmu), min_key, x_aux, w_aux, static_cast<long long>(K), mode & 1));
  g_launches.fetch_add(1);
  return SOM_OK;
}

int som_neighbourhood(const int64_t* bmu, const float* grid_pos, int64_t B, int64_t K, int64_t k_offset,
                      const float* T_dev, float* w, int64_t ldw, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!bmu || !grid_pos || !T_dev || !w || B <= 0 || K <= 0 || ldw < K)
    return fail(SOM_ERR_ARG, "som_neighbourhood: bad argument");
  dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((K + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK));
  neighbourhood_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(bmu), grid_pos, K,
                                                           k_offset, T_dev, w, ldw);
  SOM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SOM_OK;
}

int64_t som_loss_scratch_floats(int64_t B, int64_t K) {
  return B * ((K + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK) + 2;
}

int som_weighted_loss(const float* dist, int64_t ldd, const int64_t* bmu, const float* grid_pos, int64_t B, int64_t K,
                      int64_t k_offset, const float* T_dev, float inv_count, float* partials, float* loss_out,
                      void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!dist || !bmu || !grid_pos || !T_dev || !partials || !loss_out || B <= 0 || K <= 0 || ldd < K)
    return fail(SOM_ERR_ARG, "som_weighted_loss: bad argument");
  dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((K + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK));
  weighted_loss_kernel<<<grid, 256, 0, as_stream(stream)>>>(dist, ldd, reinterpret_cast<const long long*>(bmu),
                                                           grid_pos, K, k_offset, T_dev, inv_count, partials, loss_out);
  SOM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SOM_OK;
}

int som_weighted_loss_grad(const int64_t* bmu, const float* grid_pos, int64_t B, int64_t K, int64_t k_offset,
                           const float* T_dev, const float* g_out_dev, float inv_count, float* G, int64_t ldg,
                           void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!bmu || !grid_pos || !T_dev || !g_out_dev || !G || B <= 0 || K <= 0 || ldg < K)
    return fail(SOM_ERR_ARG, "som_weighted_loss_grad: bad argument");
  dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((K + COLS_PER_BLOCK - 1) / COLS_PER_BLOCK));
  loss_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(bmu), grid_pos, K, k_offset,
                                                       T_dev, g_out_dev, inv_count, G, ldg);
  SOM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SOM_OK;
}

int som_bwd_coeffs(const float* G, int64_t ldg, const float* dist, int64_t ldd, int64_t B, int64_t K, int mode,
                   const float* x_aux, const float* w_aux, float* r_hi, float* r_lo, int64_t ldr, float* ax, float* bx,
                   float* aw, float* bw, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!G || !dist || !r_hi || !r_lo || !ax || !bx || !aw || !bw || B <= 0 || K <= 0 || ldg < K || ldd < K || ldr < K)
    return fail(SOM_ERR_ARG, "som_bwd_coeffs: bad argument");
  if (mode == SOM_MODE_COSINE && (!x_aux || !w_aux)) return fail(SOM_ERR_ARG, "som_bwd_coeffs: cosine needs the reciprocal norms");
  if (mode != SOM_MODE_EUCLIDEAN && mode != SOM_MODE_COSINE) return fail(SOM_ERR_ARG, "som_bwd_coeffs: bad mode");
  dim3 grid(static_cast<unsigned>((B + COEFF_ROWS - 1) / COEFF_ROWS), static_cast<unsigned>((K + 255) / 256));
  bwd_coeffs_kernel<<<grid, 256, 0, as_stream(stream)>>>(G, ldg, dist, ldd, B, K, mode, x_aux, w_aux, r_hi, r_lo, ldr,
                                                        ax, bx, aw, bw);
  SOM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SOM_OK;
}

int som_bwd_dx(const float* r_hi, const float* r_lo, int64_t ldr, const float* w_hi, const float* w_lo, int64_t ldw,
               const float* x, int64_t ldx, const float* ax, const float* bx, int64_t B, int64_t K, int64_t D,
               float* dx, int64_t lddx, float* ws, int64_t ws_floats, void* stream) {
  if (!x || !ax || !bx || !dx || ldx < D || lddx < D) return fail(SOM_ERR_ARG, "som_bwd_dx: bad argument");
  som::EpiParams e{};
  e.alpha = ax; e.beta = bx; e.src = x; e.lds = ldx; e.out = dx; e.ldo = lddx;
  // C[B,D] = R[B,K] . W[K,D]: A = R K-major (reduction K contiguous), B = W MN-major (D contiguous)
  return launch_gemm(som::EPI_GRAD, r_hi, r_lo, ldr, 0, w_hi, w_lo, ldw, 1, B, D, K, 0, 0, 3, e, ws, ws_floats,
                     as_stream(stream));
}

int som_bwd_dw(const float* r_hi, const float* r_lo, int64_t ldr, const float* x_hi, const float* x_lo, int64_t ldx,
               const float* w, int64_t ldw, const float* aw, const float* bw, int64_t B, int64_t K, int64_t D,
               float* dw, int64_t lddw, float* ws, int64_t ws_floats, void* stream) {
  if (!w || !aw || !bw || !dw || ldw < D || lddw < D) return fail(SOM_ERR_ARG, "som_bwd_dw: bad argument");
  som::EpiParams e{};
  e.alpha = aw; e.beta = bw; e.src = w; e.lds = ldw; e.out = dw; e.ldo = lddw;
  // C[K,D] = R^T[K,B] . x[B,D]: A = R read MN-major (K contiguous), B = x MN-major (D contiguous)
  return launch_gemm(som::EPI_GRAD, r_hi, r_lo, ldr, 1, x_hi, x_lo, ldx, 1, K, D, B, 0, 0, 3, e, ws, ws_floats,
                     as_stream(stream));
}

// ---- fused protocol entry points (one call per stage of the reference's call sequence) ----------------------

int som_forward(const float* x, int64_t ldx, const float* W, int64_t ldw, int64_t B, int64_t K, int64_t D, int mode,
                int stage_w, int64_t idx_offset, float* x_hi, float* x_lo, float* x_aux, float* w_hi, float* w_lo,
                float* w_aux, int64_t ld_stage, float* dist, int64_t ldd, long long* packed, int64_t* bmu,
                int64_t K_total, float* ws, int64_t ws_floats, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!x || !x_hi || !x_lo || !x_aux || !w_hi || !w_lo || !w_aux || !packed)
    return fail(SOM_ERR_ARG, "som_forward: null pointer");
  if (stage_w && !W) return fail(SOM_ERR_ARG, "som_forward: prototypes required when stage_w is set");
  if (B <= 0 || K <= 0 || D <= 0 || D > (1ll << 30) || ldx < D || (stage_w && ldw < D) || ld_stage < D)
    return fail(SOM_ERR_ARG, "som_forward: bad shape");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_forward: bad mode");
  if ((ld_stage & (mode_f16(mode) ? 7 : 3)) != 0) return fail(SOM_ERR_ARG, "som_forward: the staging row pitch must be a multiple of 16 bytes");
  PrepSet a{x, B, ldx, x_hi, x_lo, x_aux, packed}, b{};
  if (stage_w) b = PrepSet{W, K, ldw, w_hi, w_lo, w_aux, nullptr};
  if (int rc = launch_prep(a, b, D, mode, ld_stage, as_stream(stream))) return rc;
  if (int rc = som_fwd_distances(x_hi, x_lo, ld_stage, x_aux, w_hi, w_lo, ld_stage, w_aux, B, K, D, mode, idx_offset,
                                 dist, ldd, packed, ws, ws_floats, stream))
    return rc;
  if (bmu && mode_f16(mode))
    return som_bmu_decode_scaled(packed, B, K_total > 0 ? K_total : K, bmu, nullptr, x_aux, w_aux, K, mode, stream);
  if (bmu) return som_bmu_decode(packed, B, K_total > 0 ? K_total : K, bmu, nullptr, stream);
  return SOM_OK;
}

int som_loss_fused_parts(int64_t B, int64_t K, int64_t* n_row_parts, int64_t* n_col_parts) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (B <= 0 || K <= 0 || !n_row_parts || !n_col_parts) return fail(SOM_ERR_ARG, "som_loss_fused_parts: bad argument");
  const int rpb = loss_rows_per_block(B, K, di.sms);
  *n_row_parts = (K + 127) / 128;
  *n_col_parts = (B + rpb - 1) / rpb;
  return SOM_OK;
}

int som_loss_fused(const float* dist, int64_t ldd, const int64_t* bmu, const float* grid_pos, int grid_rows,
                   int grid_cols, int64_t B, int64_t K, int64_t k_offset, const float* T_dev, float inv_count, int mode,
                   float* r_hi, float* r_lo, int64_t ldr, float* row_part, float* col_part, float* scratch,
                   float* loss_out, const float* x_aux, const float* w_aux, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!dist || !bmu || !grid_pos || !T_dev || !scratch || !loss_out || B <= 0 || K <= 0 || ldd < K)
    return fail(SOM_ERR_ARG, "som_loss_fused: bad argument");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_loss_fused: bad mode");
  const int f16 = (mode_f16(mode) && r_hi) ? 1 : 0;     // the precision only concerns the backward staging
  const int dmode = mode & 1;
  if (r_hi && (!r_lo || !row_part || !col_part || ldr < K)) return fail(SOM_ERR_ARG, "som_loss_fused: bad backward staging");
  if (f16 && (!x_aux || !w_aux || (ldr & 7) != 0 ||
              ((reinterpret_cast<uintptr_t>(r_hi) | reinterpret_cast<uintptr_t>(r_lo)) & 15) != 0))
    return fail(SOM_ERR_ARG, "som_loss_fused: fp16 staging needs the aux vectors of both stagings, ldr % 8 == 0 and aligned R");
  const float* xa = f16 ? x_aux : nullptr;
  const float* wa = f16 ? w_aux : nullptr;
  // grid_rows / grid_cols > 0: the caller vouches that grid_pos is the canonical square grid (cell k at (k / cols, k % cols))
  const bool square = grid_rows > 0 && grid_cols > 0 && grid_rows <= LC_MAX_TAB && grid_cols <= LC_MAX_TAB;
  const int rpb = loss_rows_per_block(B, K, di.sms);
  const int n_row_parts = static_cast<int>((K + 127) / 128);
  dim3 grid(static_cast<unsigned>((B + rpb - 1) / rpb), static_cast<unsigned>((K + LC_COLS - 1) / LC_COLS));
  const long long* bmu_ll = reinterpret_cast<const long long*>(bmu);
  // straight-line fast path: square grid, backward staging, everything 16-byte aligned, 4 | K, 4 | grid columns, 4 | k_offset
  const bool fast = square && r_hi && g_loss_fast.load() && (K & 3) == 0 && (grid_cols & 3) == 0 && (k_offset & 3) == 0 &&
                    (ldd & 3) == 0 && (ldr & 3) == 0 && grid_rows < 65536 && grid_cols < 65536 &&
                    ((reinterpret_cast<uintptr_t>(dist) | reinterpret_cast<uintptr_t>(r_hi) | reinterpret_cast<uintptr_t>(r_lo) |
                      reinterpret_cast<uintptr_t>(col_part)) & 15) == 0 &&
                    (!f16 || (reinterpret_cast<uintptr_t>(w_aux + K) & 15) == 0);
  auto launch_fast = [&](auto kernel) {
    return launch_kernel(kernel, grid, dim3(256), 0, as_stream(stream), dist, static_cast<long long>(ldd), bmu_ll, grid_rows,
                         grid_cols, static_cast<long long>(B), static_cast<long long>(K), static_cast<long long>(k_offset),
                         T_dev, inv_count, r_hi, r_lo, static_cast<long long>(ldr), row_part, n_row_parts, col_part, scratch,
                         loss_out, rpb, xa, wa);
  };
  if (fast) {
    if (dmode == SOM_MODE_EUCLIDEAN) {
      if (f16) SOM_CUDA(launch_fast(loss_coeffs_fast_kernel<0, true>));
      else     SOM_CUDA(launch_fast(loss_coeffs_fast_kernel<0, false>));
    } else {
      if (f16) SOM_CUDA(launch_fast(loss_coeffs_fast_kernel<1, true>));
      else     SOM_CUDA(launch_fast(loss_coeffs_fast_kernel<1, false>));
    }
  } else if (square)
    SOM_CUDA(launch_kernel(loss_coeffs_kernel<true>, grid, dim3(256), 0, as_stream(stream), dist, static_cast<long long>(ldd),
                           bmu_ll, grid_pos, grid_rows, grid_cols, static_cast<long long>(B), static_cast<long long>(K),
                           static_cast<long long>(k_offset), T_dev, inv_count, dmode, r_hi, r_lo, static_cast<long long>(ldr),
                           row_part, n_row_parts, col_part, scratch, loss_out, rpb, xa, wa));
  else
    SOM_CUDA(launch_kernel(loss_coeffs_kernel<false>, grid, dim3(256), 0, as_stream(stream), dist, static_cast<long long>(ldd),
                           bmu_ll, grid_pos, 0, 0, static_cast<long long>(B), static_cast<long long>(K),
                           static_cast<long long>(k_offset), T_dev, inv_count, dmode, r_hi, r_lo, static_cast<long long>(ldr),
                           row_part, n_row_parts, col_part, scratch, loss_out, rpb, xa, wa));
  g_launches.fetch_add(1);
  return SOM_OK;
}

int64_t som_loss_fused_scratch_floats(int64_t B, int64_t K) {
  return ((B + LC_ROWS - 1) / LC_ROWS) * ((K + LC_COLS - 1) / LC_COLS) + 2;     // bound for any rows-per-block >= LC_ROWS
}

int som_backward_dw(const float* r_hi, const float* r_lo, int64_t ldr, const float* x_hi, const float* x_lo,
                    int64_t ld_stage, const float* W, int64_t ldw, const float* col_part, int64_t n_col_parts,
                    const float* w_aux, const float* g_dev, int64_t B, int64_t K, int64_t D, int mode, float* dW,
                    int64_t lddw, int accumulate, int sm_limit, float* ws, int64_t ws_floats, void* stream) {
  if (!W || !col_part || n_col_parts <= 0 || !g_dev || !dW || ldw < D || lddw < D)
    return fail(SOM_ERR_ARG, "som_backward_dw: bad argument");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_backward_dw: bad mode");
  const int f16 = mode_f16(mode), dmode = mode & 1;
  if ((dmode == SOM_MODE_COSINE || f16) && !w_aux) return fail(SOM_ERR_ARG, "som_backward_dw: the prototype staging's aux vector is required");
  som::EpiParams e{};
  e.sum = col_part; e.sum_n = static_cast<int>(n_col_parts); e.sum_ld_m = 1; e.sum_ld_j = K;
  e.aux = w_aux; e.g_dev = g_dev; e.mode = dmode; e.accumulate = accumulate;
  e.src = W; e.lds = ldw; e.out = dW; e.ldo = lddw;
  if (f16) { e.grad_scale = w_aux + 2 * K; e.grad_inv_s = col_part + n_col_parts * K; }   // 2^g_k, 1 / S (loss kernel)
  return launch_gemm(som::EPI_GRAD, r_hi, r_lo, ldr, 1, x_hi, x_lo, ld_stage, 1, K, D, B, 0, 0, 3, e, ws, ws_floats,
                     as_stream(stream), sm_limit, f16);
}

int som_backward_dx(const float* r_hi, const float* r_lo, int64_t ldr, const float* w_hi, const float* w_lo,
                    int64_t ld_stage, const float* x, int64_t ldx, const float* row_part, int64_t n_row_parts,
                    const float* x_aux, const float* g_dev, int64_t B, int64_t K, int64_t D, int mode, float* dx,
                    int64_t lddx, int accumulate, int sm_limit, float* ws, int64_t ws_floats, void* stream) {
  if (!x || !row_part || n_row_parts <= 0 || !g_dev || !dx || ldx < D || lddx < D)
    return fail(SOM_ERR_ARG, "som_backward_dx: bad argument");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_backward_dx: bad mode");
  const int f16 = mode_f16(mode), dmode = mode & 1;
  if ((dmode == SOM_MODE_COSINE || f16) && !x_aux) return fail(SOM_ERR_ARG, "som_backward_dx: the latent staging's aux vector is required");
  som::EpiParams e{};
  e.sum = row_part; e.sum_n = static_cast<int>(n_row_parts); e.sum_ld_m = n_row_parts; e.sum_ld_j = 1;
  e.aux = x_aux; e.g_dev = g_dev; e.mode = dmode; e.accumulate = accumulate;
  e.src = x; e.lds = ldx; e.out = dx; e.ldo = lddx;
  if (f16) { e.grad_scale = x_aux + 2 * B; e.grad_inv_s = row_part + B * n_row_parts; }   // 2^e_b, 1 / S (loss kernel)
  return launch_gemm(som::EPI_GRAD, r_hi, r_lo, ldr, 0, w_hi, w_lo, ld_stage, 1, B, D, K, 0, 0, 3, e, ws, ws_floats,
                     as_stream(stream), sm_limit, f16);
}

// Both gradient GEMMs in ONE persistent CTA-pair launch: their tiles form one work list (dW tiles first) that stream-K
// spreads evenly over the CTA pairs, so there is one prologue, one tail and no tile-count quantisation per GEMM.
// Falls back to the two separate launches when the shapes do not allow a common pair tile (tiny problems) or no
// workspace was given.
// Data parallel: with dw_done != NULL every finished 32-row slab of dW adds 1 to *dw_done (release, gpu scope) and
// *dw_done_expected receives the final count, so that the exchange of dW can be started from another stream by a
// stream-ordered wait on that word (som_stream_wait_value) while the dx tiles are still being computed;
// *dw_done_expected = -1 means "not counted" (fallback launches): order the exchange after the call instead.
int som_backward_fused(const float* r_hi, const float* r_lo, int64_t ldr, const float* x_hi, const float* x_lo,
                       const float* w_hi, const float* w_lo, int64_t ld_stage, const float* x, int64_t ldx,
                       const float* W, int64_t ldw, const float* row_part, int64_t n_row_parts, const float* col_part,
                       int64_t n_col_parts, const float* x_aux, const float* w_aux, const float* g_dev, int64_t B,
                       int64_t K, int64_t D, int mode, float* dW, int64_t lddw, int accumulate_dw, float* dx,
                       int64_t lddx, int sm_limit, int count_dx, unsigned int* done, int64_t* done_expected, float* ws,
                       int64_t ws_floats, void* stream) {
  unsigned int* dw_done = done;                      // (the counted GEMM is chosen below)
  int64_t* dw_done_expected = done_expected;
  if (dw_done_expected) *dw_done_expected = -1;
  if (!W || !col_part || !x || !row_part || n_row_parts <= 0 || n_col_parts <= 0 || !g_dev || !dW || !dx || ldw < D ||
      lddw < D || ldx < D || lddx < D)
    return fail(SOM_ERR_ARG, "som_backward_fused: bad argument");
  if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_backward_fused: bad mode");
  const int mode_full = mode;
  const int f16 = mode_f16(mode);
  mode &= 1;
  if ((mode == SOM_MODE_COSINE || f16) && (!w_aux || !x_aux)) return fail(SOM_ERR_ARG, "som_backward_fused: the aux vectors of both stagings are required");
  // Counted (two-phase) launches use ALL SMs for the first GEMM and at most sm_limit for the second: the pairs beyond
  // the limit get an empty share of phase 1, exit, and their SMs are free exactly when the exchange kernel that the
  // first GEMM's completion starts needs them.  Uncounted launches: sm_limit bounds the whole grid.
  const bool phased = done != nullptr;
  int sms = 0;
  if (int rc = effective_sms(&sms, phased ? 0 : sm_limit)) return rc;
  if (ws && (reinterpret_cast<uintptr_t>(ws) & 15) != 0) return fail(SOM_ERR_ARG, "workspace must be 16-byte aligned");
  Problem p[2];
  // dW[K,D] = R^T[K,B] . x[B,D]: A = R read MN-major (K contiguous), B = x~ MN-major (D contiguous)
  p[0] = Problem{r_hi, r_lo, ldr, 1, x_hi, x_lo, ld_stage, 1, K, D, B, som::EpiParams{}};
  p[0].e.sum = col_part; p[0].e.sum_n = static_cast<int>(n_col_parts); p[0].e.sum_ld_m = 1; p[0].e.sum_ld_j = K;
  p[0].e.aux = w_aux; p[0].e.g_dev = g_dev; p[0].e.mode = mode; p[0].e.accumulate = accumulate_dw;
  p[0].e.src = W; p[0].e.lds = ldw; p[0].e.out = dW; p[0].e.ldo = lddw;
  p[0].f16 = f16;
  if (f16) { p[0].e.grad_scale = w_aux + 2 * K; p[0].e.grad_inv_s = col_part + n_col_parts * K; }
  // dx[B,D] = R[B,K] . W[K,D]: A = R K-major, B = W~ MN-major
  p[1] = Problem{r_hi, r_lo, ldr, 0, w_hi, w_lo, ld_stage, 1, B, D, K, som::EpiParams{}};
  p[1].e.sum = row_part; p[1].e.sum_n = static_cast<int>(n_row_parts); p[1].e.sum_ld_m = n_row_parts; p[1].e.sum_ld_j = 1;
  p[1].e.aux = x_aux; p[1].e.g_dev = g_dev; p[1].e.mode = mode; p[1].e.accumulate = 0;
  p[1].e.src = x; p[1].e.lds = ldx; p[1].e.out = dx; p[1].e.ldo = lddx;
  p[1].f16 = f16;
  if (f16) { p[1].e.grad_scale = x_aux + 2 * B; p[1].e.grad_inv_s = row_part + B * n_row_parts; }

  // common tile width: the cheapest stream-K schedule of the joint work list (cost model of pick_tile)
  const int forced_bn = g_bn_override.load();
  const int64_t slots = sms / 2;
  int best_bn = 0; int64_t best_workers = 0; double best_cost = 1e300;
  // Counted: two-phase schedule - every pair first works off its share of the FIRST GEMM's tiles (dW for data
  // parallelism, dx for prototype shards), then its share of the other's, so the first result is complete (and its
  // exchange can start) after about half of the launch.
  if (phased && count_dx) std::swap(p[0], p[1]);
  if (ws && g_streamk.load() >= 0 && g_cg_override.load() != 1 && B > 128 && K > 128) {
    const int pmn = som::gemm_panel_mn(f16), bk = som::gemm_bk(f16);
    for (int bn : {256, 192, 128, 64}) {
      if (forced_bn && bn != forced_bn) continue;
      if ((bn / 2) % pmn != 0) continue;              // both B operands are MN-major: whole panels per CTA
      int64_t units = 0, nkb_max = 1, units_min = INT64_MAX;
      for (int i = 0; i < 2; ++i) {
        const int64_t nkb = (p[i].Kred + bk - 1) / bk;
        const int64_t u = ((p[i].M + 255) / 256) * ((p[i].N + bn - 1) / bn) * nkb;
        units += u;
        units_min = std::min(units_min, u);
        nkb_max = std::max(nkb_max, nkb);
      }
      const int64_t workers = phased ? streamk_workers(2 * units_min, slots, bn, ws_floats, 2)
                                     : streamk_workers(units, slots, bn, ws_floats);
      if (workers < 2) continue;
      const bool panel_loads = D % pmn != 0 || !g_tma3d.load();    // B of both GEMMs has extent D
      const double cost = streamk_cost_ns(units, workers, nkb_max, bn, panel_loads);
      if (cost < best_cost) { best_cost = cost; best_bn = bn; best_workers = workers; }
    }
  }
  if (best_bn == 0) {
    if (int rc = som_backward_dw(r_hi, r_lo, ldr, x_hi, x_lo, ld_stage, W, ldw, col_part, n_col_parts, w_aux, g_dev, B, K,
                                 D, mode_full, dW, lddw, accumulate_dw, sm_limit, ws, ws_floats, stream))
      return rc;
    return som_backward_dx(r_hi, r_lo, ldr, w_hi, w_lo, ld_stage, x, ldx, row_part, n_row_parts, x_aux, g_dev, B, K, D,
                           mode_full, dx, lddx, 0, sm_limit, ws, ws_floats, stream);
  }
  int ph1 = 0;
  if (dw_done) {
    p[0].e.done_counter = dw_done;
    // 16 warp slabs (2 CTAs x 8 epilogue warps) per 256 x bn tile of the counted GEMM, each counted once by the tile's owner
    if (dw_done_expected) *dw_done_expected = ((p[0].M + 255) / 256) * ((p[0].N + best_bn - 1) / best_bn) * 16;
    if (sm_limit >= 2 && sm_limit / 2 < best_workers) ph1 = std::max(1, sm_limit / 2);
  }
  return launch_pair(som::EPI_GRAD, p, 2, best_bn, static_cast<int>(best_workers), phased ? -1 : 0,
                     pair_kchunk(g_kchunk.load(), 3), 3, ws, sms, as_stream(stream), ph1);
}

// ---- stream-ordered memory operations (driver API cuStreamWaitValue32 / cuStreamWriteValue32) ---------------------
// The data-parallel exchange of dW is enqueued on a side stream behind "wait until *addr >= value": the wait is
// executed by the GPU's front end, it occupies no SM while the gradient GEMMs (which raise the word from their
// epilogue) are still running.  Both operations can be captured in a CUDA graph.
int som_stream_wait_value(unsigned int* addr, unsigned int value, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!addr || (reinterpret_cast<uintptr_t>(addr) & 3) != 0) return fail(SOM_ERR_ARG, "som_stream_wait_value: bad address");
  static PFN_cuStreamWaitValue32_v11070 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return fail(SOM_ERR_CUDA, "cuStreamWaitValue32 entry point not available");
    fn = reinterpret_cast<PFN_cuStreamWaitValue32_v11070>(p);
  }
  const CUresult r = fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return fail(SOM_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult " + std::to_string(r));
  return SOM_OK;
}
int som_stream_write_value(unsigned int* addr, unsigned int value, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!addr || (reinterpret_cast<uintptr_t>(addr) & 3) != 0) return fail(SOM_ERR_ARG, "som_stream_write_value: bad address");
  static PFN_cuStreamWriteValue32_v11070 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return fail(SOM_ERR_CUDA, "cuStreamWriteValue32 entry point not available");
    fn = reinterpret_cast<PFN_cuStreamWriteValue32_v11070>(p);
  }
  const CUresult r = fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) return fail(SOM_ERR_CUDA, "cuStreamWriteValue32 failed with CUresult " + std::to_string(r));
  return SOM_OK;
}

// ---- prototype optimizer step fused with the operand staging of the next forward ----------------------------------
int som_adamw_step(float* W, int64_t ldw, const float* dW, int64_t lddw, float* m, float* v, int64_t ldm, int64_t K,
                   int64_t D, const float* hp_dev, double beta1, double beta2, double eps, double weight_decay, int mode,
                   float* w_hi, float* w_lo, int64_t ld_stage, float* w_aux, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!W || !dW || !m || !v || !hp_dev || K <= 0 || D <= 0 || D > (1ll << 30) || ldw < D || lddw < D || ldm < D)
    return fail(SOM_ERR_ARG, "som_adamw_step: bad argument");
  if (K > 0x7fffffffLL) return fail(SOM_ERR_ARG, "som_adamw_step: too many rows");
  if (w_hi) {
    if (!mode_ok(mode)) return fail(SOM_ERR_ARG, "som_adamw_step: bad mode");
    if (!w_lo || !w_aux || ld_stage < D || (ld_stage & (mode_f16(mode) ? 7 : 3)) != 0 ||
        ((reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo)) & 15) != 0)
      return fail(SOM_ERR_ARG, "som_adamw_step: bad staging buffers");
  }
  AdamHyper hp{beta1, beta2, eps, weight_decay};
  const int dmode = mode & 1;
  auto launch = [&](auto kernel, unsigned grid) {
    return launch_kernel(kernel, dim3(grid), dim3(256), 0, as_stream(stream), W, static_cast<long long>(ldw), dW,
                         static_cast<long long>(lddw), m, v, static_cast<long long>(ldm), static_cast<long long>(K),
                         static_cast<int>(D), hp_dev, hp, dmode, w_hi, w_lo, static_cast<long long>(ld_stage), w_aux);
  };
  const bool f16 = w_hi && mode_f16(mode);
  if (D <= 1024) {
    const unsigned grid = static_cast<unsigned>((K + 7) / 8);
    if (f16) SOM_CUDA(launch(adamw_stage_kernel<32, true>, grid));
    else     SOM_CUDA(launch(adamw_stage_kernel<32, false>, grid));
  } else {
    if (f16) SOM_CUDA(launch(adamw_stage_kernel<256, true>, static_cast<unsigned>(K)));
    else     SOM_CUDA(launch(adamw_stage_kernel<256, false>, static_cast<unsigned>(K)));
  }
  g_launches.fetch_add(1);
  return SOM_OK;
}

// In-place mean over the ranks of the fp32 buffer every rank holds at the same offset of a symmetric allocation:
//   mc_ptr    : the MULTICAST address of the buffer (NVLS; e.g. torch symmetric memory, handle.multicast_ptr)
//   flag_ptrs : device array of `world` pointers, flag_ptrs[p] = rank p's zero-initialised flag buffer of at least
//               som_nvls_flag_words(world) 32-bit words (peer-mapped symmetric allocation)
//   n_floats  : multiple of 4
// Must be launched by every rank of the group (it contains cross-GPU barriers); enqueues one kernel on `stream`.
int som_allreduce_nvls(float* mc_ptr, void* flag_ptrs, int64_t n_floats, int rank, int world, float scale, int blocks,
                       void* stream) {
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (!mc_ptr || !flag_ptrs || n_floats <= 0 || (n_floats & 3) != 0 || world < 1 || world > 32 || rank < 0 || rank >= world)
    return fail(SOM_ERR_ARG, "som_allreduce_nvls: bad argument");
  if ((reinterpret_cast<uintptr_t>(mc_ptr) & 15) != 0) return fail(SOM_ERR_ARG, "som_allreduce_nvls: misaligned buffer");
  if (blocks <= 0) blocks = 16;
  if (blocks > NVLS_MAX_BLOCKS) blocks = NVLS_MAX_BLOCKS;          // every rank must pass the same value
  nvls_allreduce_mean_kernel<<<blocks, NVLS_THREADS, 0, as_stream(stream)>>>(
      mc_ptr, reinterpret_cast<unsigned int* const*>(flag_ptrs), n_floats / 4, rank, world, scale);
  SOM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return SOM_OK;
}
int som_allreduce_mean_nvls(float* mc_ptr, void* flag_ptrs, int64_t n_floats, int rank, int world, int blocks, void* stream) {
  return som_allreduce_nvls(mc_ptr, flag_ptrs, n_floats, rank, world, 1.0f / static_cast<float>(world > 0 ? world : 1), blocks,
                            stream);
}
int64_t som_nvls_flag_words(int world) { return static_cast<int64_t>(NVLS_MAX_BLOCKS) * (world > 0 ? world : 1); }

// Host-side view of the CTA-pair kernel's work decomposition (the same sk_bound() the device code evaluates):
// bounds_out[p] = first k-block unit of worker p, bounds_out[workers] = total units.  Pure host code (tests).
int som_debug_schedule(int64_t tiles0, int64_t nkb0, int64_t tiles1, int64_t nkb1, int workers, int split,
                       int64_t* bounds_out) {
  if (tiles0 <= 0 || nkb0 <= 0 || tiles1 < 0 || (tiles1 > 0 && nkb1 <= 0) || workers <= 0 || !bounds_out)
    return fail(SOM_ERR_ARG, "som_debug_schedule: bad argument");
  som::Sched s;
  s.nprob = tiles1 > 0 ? 2 : 1;
  s.nkb0 = static_cast<int>(nkb0); s.tiles0 = static_cast<int>(tiles0);
  s.nkb1 = tiles1 > 0 ? static_cast<int>(nkb1) : 1; s.tiles1 = static_cast<int>(tiles1);
  s.units0 = tiles0 * nkb0;
  s.units = s.units0 + tiles1 * (tiles1 > 0 ? nkb1 : 0);
  s.split = split;
  if (split < 0) {                                  // two-phase schedule: phase-0 bounds, then phase-1 bounds
    if (tiles1 <= 0) return fail(SOM_ERR_ARG, "som_debug_schedule: the two-phase schedule needs two GEMMs");
    for (int ph = 0; ph < 2; ++ph)
      for (int p = 0; p <= workers; ++p) bounds_out[ph * (workers + 1) + p] = som::sk_bound_phase(s, workers, p, ph);
    return SOM_OK;
  }
  for (int p = 0; p <= workers; ++p) bounds_out[p] = som::sk_bound(s, workers, p);
  return SOM_OK;
}

int som_debug_gemm(const float* a_hi, const float* a_lo, int64_t lda, int a_mn, const float* b_hi, const float* b_lo,
                   int64_t ldb, int b_mn, int64_t M, int64_t N, int64_t Kred, int bn, int kchunk, int passes, float* C,
                   int64_t ldc, float* ws, int64_t ws_floats, void* stream) {
  if (!C || ldc < N) return fail(SOM_ERR_ARG, "som_debug_gemm: bad output");
  som::EpiParams e{};
  e.out = C; e.ldo = ldc;
  return launch_gemm(som::EPI_RAW, a_hi, a_lo, lda, a_mn, b_hi, b_lo, ldb, b_mn, M, N, Kred, bn, kchunk, passes, e,
                     ws, ws_floats, as_stream(stream));
}

// The same mainloop on fp16 operands (hi / lo point to __half matrices, lda / ldb in halves, multiples of 8).
int som_debug_gemm_f16(const void* a_hi, const void* a_lo, int64_t lda, int a_mn, const void* b_hi, const void* b_lo,
                       int64_t ldb, int b_mn, int64_t M, int64_t N, int64_t Kred, int bn, int kchunk, int passes, float* C,
                       int64_t ldc, float* ws, int64_t ws_floats, void* stream) {
  if (!C || ldc < N) return fail(SOM_ERR_ARG, "som_debug_gemm_f16: bad output");
  som::EpiParams e{};
  e.out = C; e.ldo = ldc;
  return launch_gemm(som::EPI_RAW, static_cast<const float*>(a_hi), static_cast<const float*>(a_lo), lda, a_mn,
                     static_cast<const float*>(b_hi), static_cast<const float*>(b_lo), ldb, b_mn, M, N, Kred, bn, kchunk,
                     passes, e, ws, ws_floats, as_stream(stream), 0, 1);
}

}  // extern "C"
