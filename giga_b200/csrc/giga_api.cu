// C ABI of libgiga_b200.so (see include/giga_b200.h): context, parameter packing, kernel launches.
// Pure CUDA runtime -- no torch, no CPU compute fallback.
#include "../../include/giga_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <cuda_fp16.h>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_in.cuh"
#include "decoder.cuh"
#include "decoder_ws.cuh"
#include "mise.cuh"
#include "vgn.cuh"
#include "train.cuh"
#include "train_bwd.cuh"
#include "tsdf.cuh"
#include "pack_dev.cuh"
#include "planner.cuh"
#include "unet.cuh"
#include "unet_tall.cuh"

using namespace giga;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(GIGA_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// U-Net layer tile configurations (HW, CIN0, CIN1, COUT, R, CT, TCO, CC, POOL)
using K_d0c1 = Conv3x3Cfg<40, 32, 0, 32, 4, 32, 4, 16, false>;
using K_d0c2 = Conv3x3Cfg<40, 32, 0, 32, 4, 32, 4, 16, true>;
using K_d1c1 = Conv3x3Cfg<20, 32, 0, 64, 4, 64, 4, 16, false>;
using K_d1c2 = Conv3x3Cfg<20, 64, 0, 64, 4, 64, 4, 16, true>;
using K_d2c1 = Conv3x3Cfg<10, 64, 0, 128, 10, 32, 4, 16, false>;
using K_d2c2 = Conv3x3Cfg<10, 128, 0, 128, 10, 32, 4, 16, false>;
using K_u0c1 = Conv3x3Cfg<20, 64, 64, 64, 4, 64, 4, 16, false>;
using K_u0c2 = Conv3x3Cfg<20, 64, 0, 64, 4, 64, 4, 16, false>;
using K_u1c1 = Conv3x3Cfg<40, 32, 32, 32, 4, 32, 4, 16, false>;
using K_u1c2 = Conv3x3Cfg<40, 32, 0, 32, 4, 32, 4, 16, false>;
using K_u0up = ConvTCfg<10, 128, 64, 10, 16, 32>;
using K_u1up = ConvTCfg<20, 64, 32, 4, 32, 32>;

// tensor-core U-Net on TALL pre-split activations (HW, CIN0, CIN1, COUT, NTILE, NT, MODE, FUSE_FINAL)
using T_c40 = TallCfg<40, 32, 0, 32, 32, 2, 0, false>;      // d0c1, d0c2
using T_d1c1 = TallCfg<20, 32, 0, 64, 64, 1, 0, false>;
using T_c20 = TallCfg<20, 64, 0, 64, 64, 1, 0, false>;      // d1c2, u0c2
using T_d2c1 = TallCfg<10, 64, 0, 128, 64, 1, 0, false>;
using T_d2c2 = TallCfg<10, 128, 0, 128, 64, 1, 0, false>;
using T_u0up = TallCfg<10, 128, 0, 64, 64, 1, 1, false>;
using T_u0c1 = TallCfg<20, 64, 64, 64, 64, 1, 0, false>;
using T_u1up = TallCfg<20, 64, 0, 32, 32, 2, 1, false>;
using T_u1c1 = TallCfg<40, 32, 32, 32, 32, 2, 0, false>;
using T_u1c2 = TallCfg<40, 32, 0, 32, 32, 2, 0, true>;

// persistent kernels (one CTA per SM): <cfg, stages, resident weights, N-concat>
using P_c40 = PersistCfg<T_c40, 4, true, true>;
using P_d1c1 = PersistCfg<T_d1c1, 4, false>;
using P_c20 = PersistCfg<T_c20, 4, false>;
using P_d2c1 = PersistCfg<T_d2c1, 4, false>;
using P_d2c2 = PersistCfg<T_d2c2, 4, false>;
using P_u0up = PersistCfg<T_u0up, 4, false>;
using P_u0c1 = PersistCfg<T_u0c1, 4, false>;
using P_u1up = PersistCfg<T_u1up, 4, false, true>;
using P_u1c1 = PersistCfg<T_u1c1, 3, true, true>;
using P_u1c2 = PersistCfg<T_u1c2, 4, true, true>;

struct ParamSpec {
  const char* name;
  long numel;
};

// packed encoder blob offsets (floats)
struct EncLayout {
  long conv[10];   // d0c1 d0c2 d1c1 d1c2 d2c1 d2c2 u0c1 u0c2 u1c1 u1c2 : [CIN][9][COUT]
  long bias[10];
  long up_w[2], up_b[2];
  long fin_w, fin_b;
  long tc_conv[10];  // tensor-core operand-layout weights (hi/lo fp16 splits) per conv layer
  long tc_up[2];     // transpose convs, [ab][chunk][kc][hi|lo][co][8 halfs]
  long tc_fin;       // conv_final as a 32x32 B operand [hi,lo][kc 4][n 32][8 halfs]
  long total;
};
const int kConvCin[10] = {32, 32, 32, 64, 64, 128, 128, 64, 64, 32};
const int kConvCout[10] = {32, 32, 64, 64, 128, 128, 64, 64, 32, 32};
const char* kConvName[10] = {"down_convs.0.conv1", "down_convs.0.conv2", "down_convs.1.conv1", "down_convs.1.conv2",
                             "down_convs.2.conv1", "down_convs.2.conv2", "up_convs.0.conv1",   "up_convs.0.conv2",
                             "up_convs.1.conv1",   "up_convs.1.conv2"};
const int kUpCin[2] = {128, 64}, kUpCout[2] = {64, 32};
const char* kHeadName[4] = {"qual", "rot", "width", "tsdf"};

// the encoder blob is allocated with room for the training step's data-gradient weights behind the inference layout (train_api.cuh)
long enc_blob_floats(const EncLayout& L);

EncLayout make_enc_layout() {
  EncLayout L;
  long o = 0;
  for (int i = 0; i < 10; ++i) { L.conv[i] = o; o += (long)kConvCin[i] * 9 * kConvCout[i]; L.bias[i] = o; o += kConvCout[i]; }
  for (int i = 0; i < 2; ++i) { L.up_w[i] = o; o += (long)kUpCin[i] * kUpCout[i] * 4; L.up_b[i] = o; o += kUpCout[i]; }
  L.fin_w = o; o += 32 * 32;
  L.fin_b = o; o += 32;
  // hi + lo fp16 = one 4-byte word per weight
  for (int i = 0; i < 10; ++i) { L.tc_conv[i] = o; o += (long)kConvCin[i] * kConvCout[i] * 9; }
  for (int i = 0; i < 2; ++i) { L.tc_up[i] = o; o += (long)kUpCin[i] * kUpCout[i] * 4; }
  L.tc_fin = o; o += 1024;
  L.total = o;
  return L;
}

long enc_blob_floats(const EncLayout& L) {
  long o = L.total;
  for (int i = 0; i < 10; ++i) o += (long)kConvCin[i] * 9 * kConvCout[i];
  return o;
}

// activation workspace: name -> floats per image (3B images), in forward order
struct ActSpec { const char* name; int ch, hw; };
const ActSpec kActs[] = {{"d0c1", 32, 40}, {"d0c2", 32, 40}, {"p0", 32, 20},  {"d1c1", 64, 20}, {"d1c2", 64, 20},
                         {"p1", 64, 10},   {"d2c1", 128, 10}, {"d2c2", 128, 10}, {"u0", 64, 20},  {"u0c1", 64, 20},
                         {"u0c2", 64, 20}, {"u1", 32, 40},   {"u1c1", 32, 40}, {"u1c2", 32, 40}};
constexpr int kNumActs = sizeof(kActs) / sizeof(kActs[0]);

}  // namespace

struct giga_ctx {
  int device = 0;
  std::map<std::string, std::vector<float>> raw;  // reference-layout parameters (host copies)
  bool committed = false;
  unsigned heads = 0;
  bool has_encoder = false;
  ConvInParams conv_in;
  struct TsdfMaps { const void* ptr; int B; unsigned long long use; CUtensorMap m[2]; };   // TMA tensor maps over a caller's TSDF (boxes of 7 / 3 iy rows)
  std::vector<TsdfMaps> tsdf_maps;   // small cache keyed by (pointer, B): a serving loop alternates between a few input buffers
  unsigned long long tsdf_map_use = 0;
  CUtensorMap tsdf_map[2];         // the pair selected by the last ensure_tsdf_maps()
  EncLayout el;
  float* d_enc = nullptr;    // packed encoder blob
  float* d_heads = nullptr;  // [4][DW_HEAD]   fp32 FMA-pipe decoder
  float* d_hc = nullptr;        // [4][HC_SIZE] head constants of the warp-specialised decoder (decoder_ws.cuh)
  uint8_t* d_wblob[5] = {};     // streamed weight blobs per job type (0: the three grasp heads, 1 + h: head h alone)
  unsigned* d_sched = nullptr;  // [0] next item, [1] CTAs done, [2] overflow count (self-resetting work counter)
  int decoder_impl = 1;      // 1 = warp-specialised tcgen05 3xFP16, A operand in TMEM (default), 0 = fp32 FMA pipe
  int pdl = 1;               // programmatic dependent launch between the fast-path kernels (1 = on)
  int merge_decode = 1;      // giga_forward: grasp heads + TSDF head in one decoder launch
  int tile_deps = 1;         // tile-level dependencies between consecutive same-resolution U-Net layers (needs pdl)
  unsigned long long* d_layer_times = nullptr;   // debug (GIGA_LAYER_TIMES=1): [16][4] u64
  int layer_slot = 0;
  int work_slot = 0;
  int dynamic_items = 0;     // persistent conv layers hand out items from a global counter (needs tile_deps)
  unsigned* d_flags = nullptr;   // [7 producer layers][flags_stride] LayerDep counters, zeroed every encode
  long flags_stride = 0;
  int conv_in_split = 1;     // conv_in: output channels split over this many CTAs (1 or 2) at B >= 8 (2 measured 5 % slower)
  int encoder_impl = 1;      // U-Net convs: 1 = tcgen05 3xTF32 (default), 0 = fp32 FMA pipe
  int last_impl = 0;
  int num_sms = 148;
  const char* timeline_layer = nullptr;   // debug (env GIGA_TIMELINE=<kernel name>): in-kernel phase timestamps
  unsigned long long* d_timeline = nullptr;
  long timeline_n = 0;
  // workspaces (sized for cap_B scenes)
  int cap_B = 0;
  int last_B = 0;
  float* d_pre = nullptr;      // [3][B][32][1600]
  float* d_xzpart = nullptr;   // [B][CI_NT][40][32][40]
  float* d_planes = nullptr;   // [3][B][40][40][32] plane features of giga_forward calls that pass planes = NULL
  int planes_cap = 0;
  float* d_act[kNumActs] = {};
  float* d_tall[kNumActs + 1] = {};   // TALL pre-split activations ([0] = pre, [1+i] = kActs[i]) for the tensor-core encoder
  long tall_ps[kNumActs + 1] = {};
  // host-entry staging (device side)
  struct HostSlot {   // device-side staging of one in-flight host request
    float *tsdf = nullptr, *planes = nullptr, *p = nullptr, *pt = nullptr;
    float *qual = nullptr, *rot = nullptr, *width = nullptr, *occ = nullptr;
    int cap_B = 0, cap_Ng = 0, cap_No = 0;
    cudaEvent_t ev_in = nullptr, ev_compute = nullptr, ev_done = nullptr;
    bool pending = false;
    // the slot's kernels as ONE CUDA-graph launch (captured on the second request of a configuration, replayed afterwards)
    cudaGraphExec_t graph = nullptr;
    unsigned long long graph_key = 0, graph_epoch = 0, seen_key = 0;
    long graph_launches = 0;
  };
  HostSlot slot[3];   // [0],[1]: pipelined submit/wait; [2]: the synchronous giga_forward_host
  cudaStream_t st_h2d = nullptr, st_compute = nullptr, st_d2h = nullptr;   // pipelined host path
  // planner post-processing (planner.cuh)
  int pl_cap_B = 0;
  float *d_pl_a = nullptr, *d_pl_b = nullptr, *d_pl_qlow = nullptr, *d_pl_cval = nullptr;   // [B][64000] each
  int *d_pl_cidx = nullptr, *d_pl_flag = nullptr;                                           // [B][64000], [2B] (flag | count)
  float* d_lattice = nullptr;    // [64000][3] as set by giga_ctx_set_lattice
  int det_cap_B = 0, det_cap_K = 0, det_lat_B = 0;
  float *d_det_pts = nullptr, *d_det_qual = nullptr, *d_det_rot = nullptr, *d_det_width = nullptr;   // [B][64000](x3|x4)
  float *d_det_tsdf = nullptr, *d_det_tsdfp = nullptr;                                       // host-entry staging
  int* d_det_out = nullptr;      // [B][K][4] rot | [B][K] score | [B][K] width | [B][K] index | [B] count (words)
  // giga_detect_host: pinned staging + CUDA-graph replay of the whole call (H2D, ~23 kernels, D2H)
  float* h_det_in = nullptr;     // pinned [2][B][40^3]
  int* h_det_out = nullptr;      // pinned, same layout as d_det_out
  size_t h_det_in_cap = 0, h_det_out_cap = 0;
  int use_graph = 1;
  unsigned long long graph_epoch = 0;   // bumped by anything that invalidates captured pointers / baked parameters
  struct DetGraph { unsigned long long key, epoch; cudaGraphExec_t exec; long launches; };
  std::vector<DetGraph> det_graphs;
  cudaStream_t st_graph = nullptr;
  // VGN baseline (vgn.cuh)
  bool has_vgn = false;
  float* d_vgn = nullptr;        // packed parameters
  float* d_vgn_act = nullptr;    // activations of vgn_cap_B scenes
  int vgn_cap_B = 0;
  // fused loss (train.cuh)
  float* d_loss_part = nullptr;  // [loss_cap_B][4]
  unsigned* d_loss_done = nullptr;
  int loss_cap_B = 0;
  // native training step (train_bwd.cuh): bound parameter / gradient tensors, device-packed operands, gradient workspaces
  struct Train {
    static constexpr int kSlots = 28 + 4 * 34;
    const float* val[kSlots] = {};
    float* grad[kSlots] = {};
    bool bound = false, table_dirty = true;
    int fwd_impl = 1;             // training forward: 1 = tcgen05 encoder + decoder on device-packed operands (default), 0 = fp32 FMA pipe
    unsigned heads = 0;
    PackEntry* d_tab = nullptr;
    int n_tab = 0;
    float* d_blob = nullptr;      // fp32 packed encoder operands (EncLayout's fp32 prefix) + data-gradient weights
    long dg[10] = {};             // data-gradient weights of conv i: [co][9][ci] (concat layers: two blocks, one per source)
    long blob_floats = 0;
    float* d_heads = nullptr;     // [4][DW_HEAD]
    float* d_cin = nullptr;       // conv_in [27][32] + [32]
    TcPackEntry* d_tctab = nullptr;   // device-side commit of the tensor-core operand layouts (pack_dev.cuh)
    int n_tctab = 0;
    float* d_hscale = nullptr;    // [4][2] per-head power-of-two pre-scale
    bool tc_table_dirty = true;
    int cap_B = 0;
    float* d_g[kNumActs] = {};    // gradients w.r.t. the pre-activations of kActs[i] (same shapes as d_act)
    float* d_gpre = nullptr;      // [3][B][32][1600]
    float* d_gplanes = nullptr;   // [3][B][1600][32]
    float* d_planes = nullptr;    // forward plane features
    float* d_save = nullptr;      // decoder backward scratch
    unsigned* d_wg_counters = nullptr;   // [16 launches][16 (co, ci) tiles] dynamic tile counters of the filter-gradient kernels
    int wg_slot = 0;
    size_t save_cap = 0;
    // the forward this backward belongs to
    const float *x = nullptr, *p = nullptr, *pt = nullptr;
    int B = 0, Ng = 0, No = 0, detach = 0;
    bool fwd_valid = false;
  } tr;
  // Generator3D occupancy sweep (mise.cuh): dense MISE state for one scene
  int mise_R = 0;
  unsigned char *d_mise_pstate = nullptr, *d_mise_level = nullptr, *d_mise_active = nullptr;
  float *d_mise_val = nullptr, *d_mise_qpts = nullptr, *d_mise_occ = nullptr;
  int *d_mise_qidx = nullptr, *d_mise_count = nullptr;
  long launches = 0;
  bool attrs_set = false;
  // cross-stream ordering of the shared workspaces: every public entry point that enqueues work records ev_order on its
  // stream when it returns; an entry point called on a DIFFERENT stream first makes that stream wait for the event.
  cudaEvent_t ev_order = nullptr;
  cudaStream_t order_stream = nullptr;
  bool order_valid = false, capturing = false;
  int order_depth = 0;
  size_t det_out_words = 0;
  // optional per-kernel CUDA-event timing (bench.py roofline): events recorded on the launch stream
  bool timing = false;
  struct Timed { const char* name; cudaEvent_t a, b; };
  std::vector<Timed> timed;
  std::vector<cudaEvent_t> ev_pool;
};

namespace {

int set_device(giga_ctx* ctx) {
  CU_TRY(cudaSetDevice(ctx->device));
  return GIGA_OK;
}

cudaEvent_t take_event(giga_ctx* ctx) {
  if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// Brackets one kernel launch with events when timing is on; always counts the launch.
struct LaunchScope {
  giga_ctx* ctx; cudaStream_t st; cudaEvent_t b = nullptr;
  LaunchScope(giga_ctx* c, const char* name, cudaStream_t s) : ctx(c), st(s) {
    if (ctx->timing) {
      cudaEvent_t a = take_event(ctx);
      b = take_event(ctx);
      cudaEventRecord(a, st);
      ctx->timed.push_back({name, a, b});
    }
  }
  ~LaunchScope() {
    if (b) cudaEventRecord(b, st);
    ctx->launches++;
  }
};

// One giga_ctx owns ONE set of workspaces (activations, tile-dependency flags, planner volumes): calls are serialised in
// submission order even when the caller alternates between streams (include/giga_b200.h, "Streams").  Nested entry points
// (giga_forward -> giga_encode ...) only act at the outermost level; inside a stream capture the guard is off (the capture
// stream is private to the ctx and the replay is guarded by the enclosing call).
struct OrderScope {
  giga_ctx* ctx; cudaStream_t st;
  OrderScope(giga_ctx* c, cudaStream_t s) : ctx(c), st(s) {
    if (ctx->order_depth++ == 0 && !ctx->capturing && ctx->order_valid && ctx->order_stream != st) cudaStreamWaitEvent(st, ctx->ev_order, 0);
  }
  ~OrderScope() {
    if (--ctx->order_depth == 0 && !ctx->capturing) {
      if (!ctx->ev_order) cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming);
      if (ctx->ev_order && cudaEventRecord(ctx->ev_order, st) == cudaSuccess) { ctx->order_stream = st; ctx->order_valid = true; }
    }
  }
};

// Launch on the fast path: with ctx->pdl the kernel carries the programmatic-stream-serialization attribute, i.e. it may
// start (up to its griddepcontrol.wait) while the previous kernel in the stream is still running (common.cuh).
template <typename... KArgs, typename... Args>
void launch_k(const giga_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

int ensure_attrs(giga_ctx* ctx) {
  if (ctx->attrs_set) return GIGA_OK;
  CU_TRY(cudaFuncSetAttribute(conv_in_planes_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvInCfg<5, 1>::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(conv_in_planes_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvInCfg<5, 2>::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(decode_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(decode_points_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(sample_feature_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM_BYTES));
#define SET_CONV(K) CU_TRY(cudaFuncSetAttribute(conv3x3_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES))
  SET_CONV(K_d0c1); SET_CONV(K_d0c2); SET_CONV(K_d1c1); SET_CONV(K_d1c2); SET_CONV(K_d2c1);
  SET_CONV(K_d2c2); SET_CONV(K_u0c1); SET_CONV(K_u0c2); SET_CONV(K_u1c1); SET_CONV(K_u1c2);
#undef SET_CONV
#define SET_P(P) CU_TRY(cudaFuncSetAttribute(conv_tall_persistent_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES))
  SET_P(P_c40); SET_P(P_d1c1); SET_P(P_c20); SET_P(P_d2c1); SET_P(P_d2c2); SET_P(P_u0up); SET_P(P_u0c1); SET_P(P_u1up); SET_P(P_u1c1);
  SET_P(P_u1c2);
#undef SET_P
  CU_TRY(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device));
  CU_TRY(cudaFuncSetAttribute(convT2x2_kernel<K_u0up>, cudaFuncAttributeMaxDynamicSharedMemorySize, K_u0up::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(convT2x2_kernel<K_u1up>, cudaFuncAttributeMaxDynamicSharedMemorySize, K_u1up::SMEM_BYTES));
  ctx->attrs_set = true;
  return GIGA_OK;
}

// TMA tensor maps over the caller's TSDF x[b][ix][iy][iz] (fp32): dims (innermost first) {40 iz, 40 iy, 40 ix, B}, box {48, TY + 2, 1, 1},
// out-of-volume elements zero-filled (= Conv3d's padding).  cuTensorMapEncodeTiled is a host-only driver function (no device work, ~1 us);
// it is fetched through the runtime so that the library does not link against libcuda.  Re-encoded only when (pointer, B) change.
int ensure_tsdf_maps(giga_ctx* ctx, const float* tsdf, int B) {
  for (auto& e : ctx->tsdf_maps)
    if (e.ptr == tsdf && e.B == B) {
      e.use = ++ctx->tsdf_map_use;
      ctx->tsdf_map[0] = e.m[0];
      ctx->tsdf_map[1] = e.m[1];
      return GIGA_OK;
    }
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return fail(GIGA_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  if (reinterpret_cast<uintptr_t>(tsdf) & 15) return fail(GIGA_EINVAL, "giga_encode: the TSDF pointer must be 16-byte aligned (TMA)");
  giga_ctx::TsdfMaps e;
  e.ptr = tsdf; e.B = B; e.use = ++ctx->tsdf_map_use;
  const cuuint64_t dims[4] = {(cuuint64_t)G, (cuuint64_t)G, (cuuint64_t)G, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)G * 4, (cuuint64_t)G2 * 4, (cuuint64_t)G3 * 4};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const int ty[2] = {5, 1};
  for (int i = 0; i < 2; ++i) {
    const cuuint32_t box[4] = {(cuuint32_t)CI_SLAB_W, (cuuint32_t)(ty[i] + 2), 1, 1};
    const CUresult r = encode(&e.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(tsdf), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GIGA_ECUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  }
  if (ctx->tsdf_maps.size() >= 8) {   // evict the least recently used pair
    size_t lru = 0;
    for (size_t i = 1; i < ctx->tsdf_maps.size(); ++i)
      if (ctx->tsdf_maps[i].use < ctx->tsdf_maps[lru].use) lru = i;
    ctx->tsdf_maps[lru] = e;
  } else {
    ctx->tsdf_maps.push_back(e);
  }
  ctx->tsdf_map[0] = e.m[0];
  ctx->tsdf_map[1] = e.m[1];
  return GIGA_OK;   // (captured graphs bake the map by value and are keyed on their own fixed device buffers: no epoch bump)
}

int ensure_workspace(giga_ctx* ctx, int B) {
  if (B <= ctx->cap_B) return GIGA_OK;
  ctx->graph_epoch++;
  ctx->tr.fwd_valid = false;   // the kept activations of a training forward are about to be freed
  CU_TRY(cudaDeviceSynchronize());
  if (ctx->d_pre) cudaFree(ctx->d_pre);
  if (ctx->d_xzpart) cudaFree(ctx->d_xzpart);
  for (auto& p : ctx->d_act) { if (p) cudaFree(p); p = nullptr; }
  ctx->d_pre = ctx->d_xzpart = nullptr;
  ctx->cap_B = 0;
  CU_TRY(cudaMalloc(&ctx->d_pre, sizeof(float) * 3 * (size_t)B * C * G2));
  {
    size_t units = (size_t)B * (G / 5);   // partial slabs: a later call with a smaller batch may use the finer tiling
    for (int bb = 1; bb <= B; ++bb) units = std::max(units, (size_t)bb * (G / conv_in_ty(bb)));
    CU_TRY(cudaMalloc(&ctx->d_xzpart, sizeof(float) * units * G * C * G));
  }
  for (int i = 0; i < kNumActs; ++i)
    CU_TRY(cudaMalloc(&ctx->d_act[i], sizeof(float) * 3 * (size_t)B * kActs[i].ch * kActs[i].hw * kActs[i].hw));
  for (int i = 0; i <= kNumActs; ++i) {   // zero-initialised once: padding positions are never written
    const int hw = i == 0 ? 40 : kActs[i - 1].hw, ch = i == 0 ? 32 : kActs[i - 1].ch;
    if (ctx->d_tall[i]) cudaFree(ctx->d_tall[i]);
    ctx->d_tall[i] = nullptr;
    ctx->tall_ps[i] = tall_ps(hw, 3 * B);
    const size_t bytes = sizeof(float) * (size_t)tall_floats(hw, 3 * B, ch);
    CU_TRY(cudaMalloc(&ctx->d_tall[i], bytes));
    CU_TRY(cudaMemset(ctx->d_tall[i], 0, bytes));
  }
  if (ctx->d_flags) cudaFree(ctx->d_flags);
  ctx->d_flags = nullptr;
  ctx->flags_stride = (tall_positions(40, 3 * B) + 127) / 128 + 8;   // >= groups of any layer (the 40^2 ones have the most positions)
  CU_TRY(cudaMalloc(&ctx->d_flags, sizeof(unsigned) * (9 * ctx->flags_stride + 16)));   // + 16 per-layer work counters
  CU_TRY(cudaMemset(ctx->d_flags, 0, sizeof(unsigned) * (9 * ctx->flags_stride + 16)));
  ctx->cap_B = B;
  return GIGA_OK;
}

float* act(giga_ctx* ctx, const char* name) {
  for (int i = 0; i < kNumActs; ++i)
    if (!strcmp(kActs[i].name, name)) return ctx->d_act[i];
  return nullptr;
}

template <class K>
void launch_conv(giga_ctx* ctx, const char* name, int n_img, const float* s0, const float* s1, int layer, float* out,
                 float* pooled, cudaStream_t st) {
  dim3 grid(K::NB * K::NCT, n_img);
  LaunchScope ls(ctx, name, st);
  conv3x3_kernel<K><<<grid, K::NTHREADS, K::SMEM_BYTES, st>>>(s0, s1, ctx->d_enc + ctx->el.conv[layer],
                                                              ctx->d_enc + ctx->el.bias[layer], out, pooled);
}

struct TallBuf { float* p = nullptr; long ps = 0; };

// dep_out / dep_in: index of the LayerDep counter array this layer publishes to / waits on (-1: none -> whole-grid wait).
// PIN = the producer's configuration (its group size and N-tile count define the counters' meaning).
template <class P, class PIN = P>
void launch_persist(giga_ctx* ctx, const char* name, int n_img, const TallBuf& s0, const TallBuf& s1, const float* w,
                    const float* bias, const TallBuf& out, float* fin_out, cudaStream_t st, int dep_out = -1, int dep_in = -1, int reverse = 0) {
  using K = typename P::K;
  const int n_groups = K::num_ctas(n_img), n_items = n_groups * K::NNT;
  LayerDep dep = {nullptr, nullptr, 1, 0, 0u, 0, 0, 0, nullptr, nullptr};
  if (ctx->d_layer_times) {   // debug: per-layer CTA start / end envelope (slot = launch order within the encode)
    dep.dbg = ctx->d_layer_times + 4 * (ctx->layer_slot % 16);
    ctx->layer_slot++;
  }
  dep.early_trigger = n_items >= ctx->num_sms ? 1 : 0;   // one CTA on every SM (each holds all 512 TMEM columns)
  if (ctx->tile_deps && ctx->pdl) {
    if (ctx->dynamic_items) dep.work = ctx->d_flags + 9 * ctx->flags_stride + (ctx->work_slot++ % 16);   // zeroed with the flags at the start of the encode
    if (dep_out >= 0) dep.ready_out = ctx->d_flags + (size_t)dep_out * ctx->flags_stride;
    if (dep_in >= 0) {
      using KI = typename PIN::K;
      dep.ready_in = ctx->d_flags + (size_t)dep_in * ctx->flags_stride;
      dep.in_mt = KI::MT;
      dep.in_groups = KI::num_ctas(n_img);
      dep.in_target = (unsigned)KI::NNT;
      dep.in_half_res = (KI::MODE == 1 && KI::HW * 2 == K::HW) ? 1 : 0;   // transpose conv one level down -> this concat layer
      dep.reverse = reverse;   // (kept for experiments; the hardware hands out CTAs in blockIdx order as SMs free up, so the natural
                               // order already gives the early-freed SMs the CTAs with one item more)
    }
  }
  const int grid = n_items < ctx->num_sms ? n_items : ctx->num_sms;
  unsigned long long* tl = nullptr;
  if (ctx->timeline_layer && !strcmp(ctx->timeline_layer, name)) {   // debug: per-CTA stall accounting of one layer
    const size_t n = (size_t)grid * 32;
    if (ctx->d_timeline) cudaFree(ctx->d_timeline);
    cudaMalloc(&ctx->d_timeline, n * 8);
    cudaMemsetAsync(ctx->d_timeline, 0, n * 8, st);
    ctx->timeline_n = (long)n;
    tl = ctx->d_timeline;
  }
  LaunchScope ls(ctx, name, st);
  launch_k(ctx, conv_tall_persistent_kernel<P>, dim3(grid), dim3(P::NTHREADS), P::SMEM_BYTES, st, s0.p, s0.ps, s1.p, s1.ps, w, bias, out.p,
           out.ps, ctx->d_enc + ctx->el.tc_fin, ctx->d_enc + ctx->el.fin_b, fin_out, n_img, n_groups, dep, tl);
}

template <class K>
void launch_convT(giga_ctx* ctx, const char* name, int n_img, const float* src, int up, float* out, cudaStream_t st) {
  dim3 grid(K::NB * K::NCT, n_img);
  LaunchScope ls(ctx, name, st);
  convT2x2_kernel<K><<<grid, K::NTHREADS, K::SMEM_BYTES, st>>>(src, ctx->d_enc + ctx->el.up_w[up],
                                                               ctx->d_enc + ctx->el.up_b[up], out);
}

// host float <-> half, round-to-nearest-even, bit-identical to __float2half_rn / __half2float (checked over 40 M values incl. all
// half-way cases) but inlinable: parameter commits convert 1.1 M weights and a training loop commits every step
static inline uint16_t f2h_rne(float ff) {
  uint32_t u;
  memcpy(&u, &ff, 4);
  const uint32_t f32infty = 255u << 23, f16max = (127u + 16u) << 23, denorm_magic_u = ((127u - 15u) + (23u - 10u) + 1u) << 23;
  const uint32_t sign = u & 0x80000000u;
  u ^= sign;
  uint16_t o;
  if (u >= f16max) o = (u > f32infty) ? 0x7e00 : 0x7c00;
  else if (u < (113u << 23)) {
    float f, dm;
    memcpy(&f, &u, 4); memcpy(&dm, &denorm_magic_u, 4);
    f += dm;
    memcpy(&u, &f, 4);
    o = (uint16_t)(u - denorm_magic_u);
  } else {
    const uint32_t mant_odd = (u >> 13) & 1;
    u += ((uint32_t)(15 - 127) << 23) + 0xfff;
    u += mant_odd;
    o = (uint16_t)(u >> 13);
  }
  return o | (uint16_t)(sign >> 16);
}
static inline float h2f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000) << 16, em = h & 0x7fff;
  uint32_t u;
  if (em >= 0x7c00) u = sign | 0x7f800000u | ((em & 0x3ff) << 13);
  else if (em >= 0x0400) u = sign | ((em + ((127 - 15) << 10)) << 13);
  else { float f = (float)em * (1.0f / 16777216.0f); memcpy(&u, &f, 4); u |= sign; }
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// fp16 hi / scaled lo split of a weight (same split as unet_tall.cuh::split8 for lo_scale = 2^11)
inline void split_half_host(float v, uint16_t& hi, uint16_t& lo, float lo_scale) {
  hi = f2h_rne(v);
  lo = f2h_rne((v - h2f(hi)) * lo_scale);
}

// [ntile][chunk 16 ch][tap][kc 2][hi,lo][n NTILE][8 halfs]; MODE 0 from the reference's Conv2d [co][ci][3][3],
// MODE 1 from ConvTranspose2d [ci][co][2][2] with ntile = a*2+b
template <class K>
void pack_conv_tc(const float* w, float* dst_words) {
  uint16_t* dst = reinterpret_cast<uint16_t*>(dst_words);
  for (int nt = 0; nt < K::NNT; ++nt)
    for (int c = 0; c < K::NC; ++c)
      for (int tap = 0; tap < K::NTAPS; ++tap)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < K::NTILE; ++n)
            for (int j = 0; j < 8; ++j) {
              const int ci = c * 16 + kc * 8 + j;
              float v;
              if (K::MODE == 0) v = w[((long)(nt * K::NTILE + n) * K::CIN + ci) * 9 + tap];
              else v = w[((long)ci * K::COUT + n) * 4 + nt];
              uint16_t hi, lo;
              split_half_host(v, hi, lo, LO_SCALE);
              const long base = ((((long)nt * K::NC + c) * K::NTAPS + tap) * 2) * 2 * K::NTILE * 8;
              dst[base + ((kc * 2 + 0) * K::NTILE + n) * 8 + j] = hi;     // [kc][hi|lo][n][8]
              dst[base + ((kc * 2 + 1) * K::NTILE + n) * 8 + j] = lo;
            }
}

bool get(const giga_ctx* ctx, const std::string& name, long numel, const float** out) {
  auto it = ctx->raw.find(name);
  if (it == ctx->raw.end() || (long)it->second.size() != numel) return false;
  *out = it->second.data();
  return true;
}

// tensor-core U-Net on TALL pre-split activations (unet_tall.cuh), persistent kernels: d_tall[0] (pre) -> planes (conv_final fused into the
// last layer), or -> d_tall["u1c2"] when keep_u1c2
void unet_tall_forward(giga_ctx* ctx, int n_img, float* planes, bool keep_u1c2, cudaStream_t st) {
    auto tb = [&](const char* nm) {
      TallBuf t;
      if (!strcmp(nm, "pre")) { t.p = ctx->d_tall[0]; t.ps = ctx->tall_ps[0]; return t; }
      for (int i = 0; i < kNumActs; ++i)
        if (!strcmp(kActs[i].name, nm)) { t.p = ctx->d_tall[1 + i]; t.ps = ctx->tall_ps[1 + i]; }
      return t;
    };
    const TallBuf none;
    const float* E = ctx->d_enc;
    const EncLayout& L = ctx->el;
    {
      // tile-level dependency edges (dep_out -> dep_in): d0c1->d0c2, d1c1->d1c2, d2c1->d2c2->u0up->u0c1->u0c2->u1up->u1c1->u1c2;
      // the edges through the max-pools are whole-grid waits
      launch_persist<P_c40>(ctx, "conv3x3:d0c1", n_img, tb("pre"), none, E + L.tc_conv[0], E + L.bias[0], tb("d0c1"), nullptr, st, 0, -1);
      launch_persist<P_c40, P_c40>(ctx, "conv3x3:d0c2", n_img, tb("d0c1"), none, E + L.tc_conv[1], E + L.bias[1], tb("d0c2"), nullptr, st, -1, 0, 0);
      {
        LaunchScope ls(ctx, "maxpool:p0", st);
        launch_k(ctx, pool_tall_kernel<20, 4>, dim3(ceil_div(n_img * 4 * 400, 256)), dim3(256), 0, st, (const float*)tb("d0c2").p, tb("d0c2").ps,
                 tb("p0").p, tb("p0").ps, n_img);
      }
      launch_persist<P_d1c1>(ctx, "conv3x3:d1c1", n_img, tb("p0"), none, E + L.tc_conv[2], E + L.bias[2], tb("d1c1"), nullptr, st, 1, -1);
      launch_persist<P_c20, P_d1c1>(ctx, "conv3x3:d1c2", n_img, tb("d1c1"), none, E + L.tc_conv[3], E + L.bias[3], tb("d1c2"), nullptr, st, -1, 1, 0);
      {
        LaunchScope ls(ctx, "maxpool:p1", st);
        launch_k(ctx, pool_tall_kernel<10, 8>, dim3(ceil_div(n_img * 8 * 100, 256)), dim3(256), 0, st, (const float*)tb("d1c2").p, tb("d1c2").ps,
                 tb("p1").p, tb("p1").ps, n_img);
      }
      launch_persist<P_d2c1>(ctx, "conv3x3:d2c1", n_img, tb("p1"), none, E + L.tc_conv[4], E + L.bias[4], tb("d2c1"), nullptr, st, 2, -1);
      launch_persist<P_d2c2, P_d2c1>(ctx, "conv3x3:d2c2", n_img, tb("d2c1"), none, E + L.tc_conv[5], E + L.bias[5], tb("d2c2"), nullptr, st, 3, 2, 0);
      launch_persist<P_u0up, P_d2c2>(ctx, "convT:u0", n_img, tb("d2c2"), none, E + L.tc_up[0], E + L.up_b[0], tb("u0"), nullptr, st, 7, 3, 0);
      launch_persist<P_u0c1, P_u0up>(ctx, "conv3x3:u0c1", n_img, tb("u0"), tb("d1c2"), E + L.tc_conv[6], E + L.bias[6], tb("u0c1"), nullptr, st, 4, 7);
      launch_persist<P_c20, P_u0c1>(ctx, "conv3x3:u0c2", n_img, tb("u0c1"), none, E + L.tc_conv[7], E + L.bias[7], tb("u0c2"), nullptr, st, 5, 4, 0);
      launch_persist<P_u1up, P_c20>(ctx, "convT:u1", n_img, tb("u0c2"), none, E + L.tc_up[1], E + L.up_b[1], tb("u1"), nullptr, st, 8, 5, 0);
      launch_persist<P_u1c1, P_u1up>(ctx, "conv3x3:u1c1", n_img, tb("u1"), tb("d0c2"), E + L.tc_conv[8], E + L.bias[8], tb("u1c1"), nullptr, st, 6, 8);
      if (keep_u1c2)   // the training step needs u1c2 itself (ReLU mask, conv_final's filter gradient): plain layer, conv_final runs separately
        launch_persist<P_c40, P_u1c1>(ctx, "conv3x3:u1c2", n_img, tb("u1c1"), none, E + L.tc_conv[9], E + L.bias[9], tb("u1c2"), nullptr, st, -1, 6, 0);
      else
        launch_persist<P_u1c2, P_u1c1>(ctx, "conv3x3:u1c2+final", n_img, tb("u1c1"), none, E + L.tc_conv[9], E + L.bias[9], none, planes, st, -1, 6, 0);
    }
}

}  // namespace

extern "C" {

int giga_version(void) { return 100; }

const char* giga_last_error(void) { return g_err.c_str(); }

int giga_ctx_create(giga_ctx** out, int device) {
  if (!out) return fail(GIGA_EINVAL, "giga_ctx_create: out is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(GIGA_ENODEV, std::string("no CUDA device (") + cudaGetErrorString(e) + "); giga_b200 has no CPU path");
  if (device < 0 || device >= n) return fail(GIGA_EINVAL, "giga_ctx_create: bad device ordinal");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(GIGA_ENODEV, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                 std::to_string(prop.minor) + "; this library is built for sm_100a only");
  giga_ctx* ctx = new giga_ctx();
  ctx->device = device;
  ctx->timeline_layer = getenv("GIGA_TIMELINE");
  if (const char* e = getenv("GIGA_PDL")) ctx->pdl = atoi(e) != 0;   // A/B switches (debug)
  if (const char* e = getenv("GIGA_MERGE_DECODE")) ctx->merge_decode = atoi(e) != 0;
  if (const char* e = getenv("GIGA_TILE_DEPS")) ctx->tile_deps = atoi(e) != 0;
  if (const char* e = getenv("GIGA_DYNAMIC")) ctx->dynamic_items = atoi(e) != 0;
  if (getenv("GIGA_LAYER_TIMES")) {
    cudaMalloc(&ctx->d_layer_times, 16 * 4 * 8);
    cudaMemset(ctx->d_layer_times, 0, 16 * 4 * 8);
  }
  if (const char* e = getenv("GIGA_CONV_IN_SPLIT")) ctx->conv_in_split = atoi(e) == 2 ? 2 : 1;
  ctx->el = make_enc_layout();
  *out = ctx;
  return GIGA_OK;
}

void giga_ctx_destroy(giga_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  float* ptrs[] = {ctx->d_enc, ctx->d_heads, ctx->d_pre, ctx->d_xzpart, ctx->d_planes, reinterpret_cast<float*>(ctx->d_flags)};
  for (float* p : ptrs)
    if (p) cudaFree(p);
  for (auto& s : ctx->slot) {
    float* sp[] = {s.tsdf, s.planes, s.p, s.pt, s.qual, s.rot, s.width, s.occ};
    for (float* p : sp)
      if (p) cudaFree(p);
    if (s.graph) cudaGraphExecDestroy(s.graph);
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_compute) cudaEventDestroy(s.ev_compute);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
  }
  if (ctx->st_h2d) cudaStreamDestroy(ctx->st_h2d);
  if (ctx->st_compute) cudaStreamDestroy(ctx->st_compute);
  if (ctx->st_d2h) cudaStreamDestroy(ctx->st_d2h);
  if (ctx->d_timeline) cudaFree(ctx->d_timeline);
  {
    void* mp[] = {ctx->d_mise_pstate, ctx->d_mise_level, ctx->d_mise_active, ctx->d_mise_val, ctx->d_mise_qpts, ctx->d_mise_occ, ctx->d_mise_qidx,
                  ctx->d_mise_count};
    for (void* q : mp)
      if (q) cudaFree(q);
  }
  {
    auto& T = ctx->tr;   // (d_blob / d_heads alias ctx->d_enc / ctx->d_heads)
    void* tp[] = {T.d_tab, T.d_cin, T.d_tctab, T.d_hscale, T.d_gpre, T.d_gplanes, T.d_planes, T.d_save, T.d_wg_counters};
    for (void* q : tp)
      if (q) cudaFree(q);
    for (float* q : T.d_g)
      if (q) cudaFree(q);
  }
  if (ctx->d_vgn) cudaFree(ctx->d_vgn);
  if (ctx->d_vgn_act) cudaFree(ctx->d_vgn_act);
  if (ctx->d_loss_part) cudaFree(ctx->d_loss_part);
  if (ctx->d_loss_done) cudaFree(ctx->d_loss_done);
  if (ctx->d_hc) cudaFree(ctx->d_hc);
  if (ctx->d_sched) cudaFree(ctx->d_sched);
  if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
  for (auto& b : ctx->d_wblob)
    if (b) cudaFree(b);
  for (auto& g : ctx->det_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (ctx->st_graph) cudaStreamDestroy(ctx->st_graph);
  if (ctx->d_layer_times) cudaFree(ctx->d_layer_times);
  if (ctx->h_det_in) cudaFreeHost(ctx->h_det_in);
  if (ctx->h_det_out) cudaFreeHost(ctx->h_det_out);
  void* pl[] = {ctx->d_pl_a, ctx->d_pl_b, ctx->d_pl_qlow, ctx->d_pl_cval, ctx->d_pl_cidx, ctx->d_pl_flag, ctx->d_lattice, ctx->d_det_pts,
                ctx->d_det_qual, ctx->d_det_rot, ctx->d_det_width, ctx->d_det_tsdf, ctx->d_det_tsdfp, ctx->d_det_out};
  for (void* p : pl)
    if (p) cudaFree(p);
  for (float* p : ctx->d_act)
    if (p) cudaFree(p);
  for (float* p : ctx->d_tall)
    if (p) cudaFree(p);
  for (auto& t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  delete ctx;
}

int giga_ctx_set_param(giga_ctx* ctx, const char* name, const float* data, long numel, int on_device) {
  if (!ctx || !name || !data || numel <= 0) return fail(GIGA_EINVAL, "giga_ctx_set_param: bad argument");
  if (int r = set_device(ctx)) return r;
  std::vector<float>& v = ctx->raw[name];
  v.resize(numel);
  if (on_device) {
    CU_TRY(cudaDeviceSynchronize());   // the producer of `data` may run on a non-blocking stream the legacy-stream copy does not wait for
    CU_TRY(cudaMemcpy(v.data(), data, sizeof(float) * numel, cudaMemcpyDeviceToHost));
  } else {
    memcpy(v.data(), data, sizeof(float) * numel);
  }
  ctx->committed = false;
  return GIGA_OK;
}

int giga_ctx_set_params_flat(giga_ctx* ctx, int n, const char* const* names, const long* offsets, const long* numels, const float* flat,
                             long total, int on_device) {
  if (!ctx || n <= 0 || !names || !offsets || !numels || !flat || total <= 0) return fail(GIGA_EINVAL, "giga_ctx_set_params_flat: bad argument");
  if (int r = set_device(ctx)) return r;
  std::vector<float> host;
  const float* src = flat;
  if (on_device) {
    host.resize(total);
    CU_TRY(cudaDeviceSynchronize());   // see giga_ctx_set_param
    CU_TRY(cudaMemcpy(host.data(), flat, sizeof(float) * total, cudaMemcpyDeviceToHost));
    src = host.data();
  }
  for (int i = 0; i < n; ++i) {
    if (!names[i] || numels[i] <= 0 || offsets[i] < 0 || offsets[i] + numels[i] > total)
      return fail(GIGA_EINVAL, "giga_ctx_set_params_flat: tensor outside the flat buffer");
    std::vector<float>& v = ctx->raw[names[i]];
    v.assign(src + offsets[i], src + offsets[i] + numels[i]);
  }
  ctx->committed = false;
  return GIGA_OK;
}

int giga_ctx_commit_params(giga_ctx* ctx) {
  if (!ctx) return fail(GIGA_EINVAL, "giga_ctx_commit_params: ctx is null");
  if (int r = set_device(ctx)) return r;
  // the packed blobs below are overwritten with synchronous copies on the legacy stream, which does not order against the non-blocking
  // streams kernels of this ctx may still be running on (torch side streams, the pipelined host path): drain the device first
  CU_TRY(cudaDeviceSynchronize());
  const float* w;
  const float* b;
  // ---- encoder ----
  ctx->has_encoder = false;
  if (ctx->raw.count("encoder.conv_in.weight")) {
    if (!get(ctx, "encoder.conv_in.weight", 32 * 27, &w) || !get(ctx, "encoder.conv_in.bias", 32, &b))
      return fail(GIGA_ESTATE, "encoder.conv_in.{weight,bias}: missing or wrong size");
    for (int c = 0; c < 32; ++c) {
      for (int t = 0; t < 27; ++t) (&ctx->conv_in.w[t][0].x)[c] = w[c * 27 + t];  // [co][0][dx][dy][dz] -> [tap][co]
      (&ctx->conv_in.b[0].x)[c] = b[c];
    }
    std::vector<float> blob(ctx->el.total, 0.f);
    for (int i = 0; i < 10; ++i) {
      const int ci = kConvCin[i], co = kConvCout[i];
      std::string base = std::string("encoder.unet.") + kConvName[i];
      if (!get(ctx, base + ".weight", (long)co * ci * 9, &w) || !get(ctx, base + ".bias", co, &b))
        return fail(GIGA_ESTATE, base + ".{weight,bias}: missing or wrong size");
      float* dst = blob.data() + ctx->el.conv[i];  // [ci][tap][co]  <-  [co][ci][3][3]
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int t = 0; t < 9; ++t) dst[((long)c * 9 + t) * co + o] = w[((long)o * ci + c) * 9 + t];
      memcpy(blob.data() + ctx->el.bias[i], b, sizeof(float) * co);
    }
    for (int i = 0; i < 2; ++i) {
      std::string base = "encoder.unet.up_convs." + std::to_string(i) + ".upconv";
      if (!get(ctx, base + ".weight", (long)kUpCin[i] * kUpCout[i] * 4, &w) || !get(ctx, base + ".bias", kUpCout[i], &b))
        return fail(GIGA_ESTATE, base + ".{weight,bias}: missing or wrong size");
      memcpy(blob.data() + ctx->el.up_w[i], w, sizeof(float) * kUpCin[i] * kUpCout[i] * 4);  // [ci][co][2][2] as is
      memcpy(blob.data() + ctx->el.up_b[i], b, sizeof(float) * kUpCout[i]);
    }
    if (!get(ctx, "encoder.unet.conv_final.weight", 32 * 32, &w) || !get(ctx, "encoder.unet.conv_final.bias", 32, &b))
      return fail(GIGA_ESTATE, "encoder.unet.conv_final.{weight,bias}: missing or wrong size");
    for (int o = 0; o < 32; ++o)
      for (int c = 0; c < 32; ++c) blob[ctx->el.fin_w + c * 32 + o] = w[o * 32 + c];  // [ci][co]
    memcpy(blob.data() + ctx->el.fin_b, b, sizeof(float) * 32);
    {
      uint16_t* fw = reinterpret_cast<uint16_t*>(blob.data() + ctx->el.tc_fin);   // [hi|lo][kc 4][n 32][8 halfs]
      for (int k = 0; k < 32; ++k)
        for (int n = 0; n < 32; ++n) {
          uint16_t hi, lo;
          split_half_host(w[n * 32 + k], hi, lo, LO_SCALE);
          fw[((k / 8) * 32 + n) * 8 + (k % 8)] = hi;
          fw[1024 + ((k / 8) * 32 + n) * 8 + (k % 8)] = lo;
        }
    }
    {
      const float* cw[10];
      for (int i = 0; i < 10; ++i) get(ctx, std::string("encoder.unet.") + kConvName[i] + ".weight", (long)kConvCout[i] * kConvCin[i] * 9, &cw[i]);
      pack_conv_tc<T_c40>(cw[0], blob.data() + ctx->el.tc_conv[0]);
      pack_conv_tc<T_c40>(cw[1], blob.data() + ctx->el.tc_conv[1]);
      pack_conv_tc<T_d1c1>(cw[2], blob.data() + ctx->el.tc_conv[2]);
      pack_conv_tc<T_c20>(cw[3], blob.data() + ctx->el.tc_conv[3]);
      pack_conv_tc<T_d2c1>(cw[4], blob.data() + ctx->el.tc_conv[4]);
      pack_conv_tc<T_d2c2>(cw[5], blob.data() + ctx->el.tc_conv[5]);
      pack_conv_tc<T_u0c1>(cw[6], blob.data() + ctx->el.tc_conv[6]);
      pack_conv_tc<T_c20>(cw[7], blob.data() + ctx->el.tc_conv[7]);
      pack_conv_tc<T_u1c1>(cw[8], blob.data() + ctx->el.tc_conv[8]);
      pack_conv_tc<T_u1c2>(cw[9], blob.data() + ctx->el.tc_conv[9]);
      const float* uw;
      get(ctx, "encoder.unet.up_convs.0.upconv.weight", (long)kUpCin[0] * kUpCout[0] * 4, &uw);
      pack_conv_tc<T_u0up>(uw, blob.data() + ctx->el.tc_up[0]);
      get(ctx, "encoder.unet.up_convs.1.upconv.weight", (long)kUpCin[1] * kUpCout[1] * 4, &uw);
      pack_conv_tc<T_u1up>(uw, blob.data() + ctx->el.tc_up[1]);
    }
    if (!ctx->d_enc) CU_TRY(cudaMalloc(&ctx->d_enc, sizeof(float) * enc_blob_floats(ctx->el)));   // + room for the training step's data-gradient weights
    CU_TRY(cudaMemcpy(ctx->d_enc, blob.data(), sizeof(float) * ctx->el.total, cudaMemcpyHostToDevice));
    ctx->has_encoder = true;
  }
  // ---- decoder heads ----
  std::vector<float> hb((size_t)4 * DW_HEAD, 0.f);
  unsigned heads = 0;
  int head_sexp[4] = {0, 0, 0, 0};
  for (int h = 0; h < 4; ++h) {
    const std::string pre = std::string("decoder_") + kHeadName[h] + ".";
    if (!ctx->raw.count(pre + "fc_p.weight")) continue;
    const int od = (h == 1) ? 4 : 1;
    float* H = hb.data() + (size_t)h * DW_HEAD;
    auto need = [&](const std::string& n, long numel, const float** p) { return get(ctx, pre + n, numel, p); };
    if (!need("fc_p.weight", 96, &w) || !need("fc_p.bias", 32, &b)) return fail(GIGA_ESTATE, pre + "fc_p: missing or wrong size");
    for (int j = 0; j < 32; ++j) {
      for (int k = 0; k < 3; ++k) H[DW_FCP + k * 32 + j] = w[j * 3 + k];
      H[DW_FCP + 96 + j] = b[j];
    }
    for (int i = 0; i < 5; ++i) {
      float* Bk = H + DW_BLOCK0 + i * DW_BLK;
      const std::string si = std::to_string(i);
      if (!need("fc_c." + si + ".weight", 32 * 96, &w) || !need("fc_c." + si + ".bias", 32, &b))
        return fail(GIGA_ESTATE, pre + "fc_c." + si + ": missing or wrong size");
      for (int j = 0; j < 32; ++j) {
        for (int k = 0; k < 96; ++k) Bk[DW_BLK_FCC + k * 32 + j] = w[j * 96 + k];
        Bk[DW_BLK_BC + j] = b[j];
      }
      const int woff[2] = {DW_BLK_W0, DW_BLK_W1}, boff[2] = {DW_BLK_B0, DW_BLK_B1};
      for (int f = 0; f < 2; ++f) {
        const std::string fn = "blocks." + si + ".fc_" + std::to_string(f);
        if (!need(fn + ".weight", 32 * 32, &w) || !need(fn + ".bias", 32, &b))
          return fail(GIGA_ESTATE, pre + fn + ": missing or wrong size");
        for (int j = 0; j < 32; ++j) {
          for (int k = 0; k < 32; ++k) Bk[woff[f] + k * 32 + j] = w[j * 32 + k];
          Bk[boff[f] + j] = b[j];
        }
      }
    }
    if (!need("fc_out.weight", od * 32, &w) || !need("fc_out.bias", od, &b))
      return fail(GIGA_ESTATE, pre + "fc_out: missing or wrong size");
    for (int m = 0; m < od; ++m) {
      for (int k = 0; k < 32; ++k) H[DW_OUT + k * 4 + m] = w[m * 32 + k];
      H[DW_OUT + 128 + m] = b[m];
    }
    // per-head power-of-two weight pre-scale of the tensor-core decoder: max |w| * 2^s in [512, 1024), so that the hi/lo fp16 halves of
    // every non-negligible weight are normal fp16 numbers
    float wmax = 0.f;
    for (int i = 0; i < 5; ++i) {
      const std::string si = std::to_string(i);
      need("fc_c." + si + ".weight", 32 * 96, &w);
      for (int e = 0; e < 32 * 96; ++e) wmax = fmaxf(wmax, fabsf(w[e]));
      for (int f = 0; f < 2; ++f) {
        need("blocks." + si + ".fc_" + std::to_string(f) + ".weight", 32 * 32, &w);
        for (int e = 0; e < 32 * 32; ++e) wmax = fmaxf(wmax, fabsf(w[e]));
      }
    }
    int sexp = 0;
    if (wmax > 0.f && std::isfinite(wmax)) {
      int e;
      frexpf(wmax, &e);           // wmax = m * 2^e, m in [0.5, 1)
      sexp = 10 - e;
      if (sexp < -14) sexp = -14;
      if (sexp > 24) sexp = 24;
    }
    head_sexp[h] = sexp;
    heads |= 1u << h;
  }
  {
    // ---- warp-specialised decoder (decoder_ws.cuh): head constants + one streamed weight blob per job type ----
    std::vector<float> hc((size_t)4 * HC_SIZE, 0.f);
    for (int h = 0; h < 4; ++h) {
      if (!(heads & (1u << h))) continue;
      const float* H = hb.data() + (size_t)h * DW_HEAD;
      float* D = hc.data() + (size_t)h * HC_SIZE;
      memcpy(D + HC_FCP, H + DW_FCP, sizeof(float) * 128);
      memcpy(D + HC_OUT, H + DW_OUT, sizeof(float) * 132);
      memcpy(D + HC_B1, H + DW_BLOCK0 + 4 * DW_BLK + DW_BLK_B1, sizeof(float) * 32);
      D[HC_INV] = ldexpf(1.f, -head_sexp[h]);
    }
    if (!ctx->d_hc) CU_TRY(cudaMalloc(&ctx->d_hc, sizeof(float) * 4 * HC_SIZE));
    CU_TRY(cudaMemcpy(ctx->d_hc, hc.data(), sizeof(float) * 4 * HC_SIZE, cudaMemcpyHostToDevice));
    if (!ctx->d_sched) {
      CU_TRY(cudaMalloc(&ctx->d_sched, sizeof(unsigned) * 4));
      CU_TRY(cudaMemset(ctx->d_sched, 0, sizeof(unsigned) * 4));
    }
    for (int type = 0; type < 5; ++type) {
      const int nh = type == 0 ? 3 : 1, head0 = type == 0 ? 0 : type - 1;
      const unsigned need = type == 0 ? 7u : (1u << head0);
      if ((heads & need) != need) {
        if (ctx->d_wblob[type]) { cudaFree(ctx->d_wblob[type]); ctx->d_wblob[type] = nullptr; }
        continue;
      }
      std::vector<uint8_t> blob((size_t)5 * wd_block_bytes(nh), 0);
      for (int blk = 0; blk < 5; ++blk) {
        uint8_t* base = blob.data() + (size_t)blk * wd_block_bytes(nh);
        uint16_t* fcc = reinterpret_cast<uint16_t*>(base);                         // [hi|lo][kc 12][n nh*32][8]
        uint16_t* ch = reinterpret_cast<uint16_t*>(base + wd_fcc_bytes(nh));       // W0 [head][hi|lo][kc 4][n 32][8], W1 same
        float* bias = reinterpret_cast<float*>(base + wd_fcc_bytes(nh) + nh * 8192);   // b0 [nh][32], bin [nh][32]
        for (int c = 0; c < nh; ++c) {
          const float* Hb = hb.data() + (size_t)(head0 + c) * DW_HEAD + DW_BLOCK0 + blk * DW_BLK;   // input-major fp32 copies: Wt[k][j]
          const float wscale = ldexpf(1.f, head_sexp[head0 + c]);
          for (int k = 0; k < 96; ++k)
            for (int j = 0; j < 32; ++j) {
              uint16_t hi, lo;
              split_half_host(Hb[DW_BLK_FCC + k * 32 + j] * wscale, hi, lo, 1.f);
              const size_t e = ((size_t)(k / 8) * (nh * 32) + c * 32 + j) * 8 + (k % 8);
              fcc[e] = hi;
              fcc[(size_t)12 * nh * 32 * 8 + e] = lo;
            }
          const int woff[2] = {DW_BLK_W0, DW_BLK_W1};
          for (int f = 0; f < 2; ++f)
            for (int k = 0; k < 32; ++k)
              for (int j = 0; j < 32; ++j) {
                uint16_t hi, lo;
                split_half_host(Hb[woff[f] + k * 32 + j] * wscale, hi, lo, 1.f);
                const size_t e = (size_t)f * nh * 2048 + (size_t)c * 2048 + (k / 8) * 256 + j * 8 + (k % 8);
                ch[e] = hi;
                ch[e + 1024] = lo;
              }
          for (int j = 0; j < 32; ++j) {
            bias[c * 32 + j] = Hb[DW_BLK_B0 + j];
            bias[nh * 32 + c * 32 + j] = Hb[DW_BLK_BC + j] + (blk > 0 ? Hb[DW_BLK_B1 + j - DW_BLK] : 0.f);   // bc_b + b1_{b-1}
          }
        }
      }
      if (!ctx->d_wblob[type]) CU_TRY(cudaMalloc(&ctx->d_wblob[type], blob.size()));
      CU_TRY(cudaMemcpy(ctx->d_wblob[type], blob.data(), blob.size(), cudaMemcpyHostToDevice));
    }
  }
  if (!ctx->d_heads) CU_TRY(cudaMalloc(&ctx->d_heads, sizeof(float) * 4 * DW_HEAD));
  CU_TRY(cudaMemcpy(ctx->d_heads, hb.data(), sizeof(float) * 4 * DW_HEAD, cudaMemcpyHostToDevice));
  ctx->heads = heads;
  // ---- VGN baseline (networks.py:48-63): encoder.conv{1,2,3}, decoder.conv{1,2,3}, conv_{qual,rot,width} ----
  ctx->has_vgn = false;
  if (ctx->raw.count("conv_qual.weight")) {
    const VgnLayout VL = make_vgn_layout();
    std::vector<float> vb(VL.total, 0.f);
    const char* lname[6] = {"encoder.conv1", "encoder.conv2", "encoder.conv3", "decoder.conv1", "decoder.conv2", "decoder.conv3"};
    for (int i = 0; i < 6; ++i) {
      const int ci = kVgnCin[i], co = kVgnCout[i], k3 = kVgnK[i] * kVgnK[i] * kVgnK[i];
      if (!get(ctx, std::string(lname[i]) + ".weight", (long)co * ci * k3, &w) || !get(ctx, std::string(lname[i]) + ".bias", co, &b))
        return fail(GIGA_ESTATE, std::string(lname[i]) + ".{weight,bias}: missing or wrong size");
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int t = 0; t < k3; ++t) vb[VL.w[i] + ((long)c * k3 + t) * co + o] = w[((long)o * ci + c) * k3 + t];   // [co][ci][k^3] -> [ci][k^3][co]
      memcpy(vb.data() + VL.b[i], b, sizeof(float) * co);
    }
    const char* hname[3] = {"conv_qual", "conv_rot", "conv_width"};
    const int hch[3] = {1, 4, 1}, hoff[3] = {0, 1, 5};
    for (int h = 0; h < 3; ++h) {
      if (!get(ctx, std::string(hname[h]) + ".weight", (long)hch[h] * 16 * 125, &w) || !get(ctx, std::string(hname[h]) + ".bias", hch[h], &b))
        return fail(GIGA_ESTATE, std::string(hname[h]) + ".{weight,bias}: missing or wrong size");
      for (int o = 0; o < hch[h]; ++o) {
        for (int c = 0; c < 16; ++c)
          for (int t = 0; t < 125; ++t) vb[VL.w[6] + ((long)c * 125 + t) * 8 + hoff[h] + o] = w[((long)o * 16 + c) * 125 + t];
        vb[VL.b[6] + hoff[h] + o] = b[o];
      }
    }
    if (!ctx->d_vgn) CU_TRY(cudaMalloc(&ctx->d_vgn, sizeof(float) * VL.total));
    CU_TRY(cudaMemcpy(ctx->d_vgn, vb.data(), sizeof(float) * VL.total, cudaMemcpyHostToDevice));
    ctx->has_vgn = true;
  }
  if (!ctx->has_encoder && !heads && !ctx->has_vgn) return fail(GIGA_ESTATE, "giga_ctx_commit_params: no parameters were set");
  if (int r = ensure_attrs(ctx)) return r;
  CU_TRY(cudaDeviceSynchronize());   // pageable-source copies may still be in flight when cudaMemcpy returns
  ctx->committed = true;
  ctx->graph_epoch++;   // conv_in's weights are kernel parameters (baked into captured graphs)
  return GIGA_OK;
}

unsigned giga_ctx_heads(const giga_ctx* ctx) { return ctx ? ctx->heads : 0; }

int giga_encode(giga_ctx* ctx, const float* tsdf, int B, float* planes, void* stream) {
  if (!ctx || !tsdf || !planes || B <= 0) return fail(GIGA_EINVAL, "giga_encode: bad argument");
  if (!ctx->committed || !ctx->has_encoder) return fail(GIGA_ESTATE, "giga_encode: encoder parameters not committed");
  if (int r = set_device(ctx)) return r;
  if (int r = ensure_workspace(ctx, B)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  const int n_img = 3 * B;
  if (ctx->d_layer_times) {   // debug: (re)initialise the envelopes: min fields to ~0, max fields to 0
    static const unsigned long long init[64] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0,
                                                ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0,
                                                ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
    cudaMemcpyAsync(ctx->d_layer_times, init, sizeof init, cudaMemcpyHostToDevice, st);
    ctx->layer_slot = 0;
  }
  // ---- fused Conv3d + ReLU + tri-plane means: TSDF slabs staged by TMA (tensor map over the caller's volume), planes leave as TALL
  //      pre-split operands of the first U-Net layer ----
  if (int r = ensure_tsdf_maps(ctx, tsdf, B)) return r;
  float* tall_pre = ctx->d_tall[0];
  const long ps_pre = ctx->tall_ps[0];
  {
    LaunchScope ls(ctx, "conv_in_planes", st);
    if (conv_in_ty(B) == 1) {
      using Cf = ConvInCfg<1, 4>;
      launch_k(ctx, conv_in_planes_kernel<1, 4>, dim3(Cf::NT, B, 4), dim3(Cf::THREADS), Cf::SMEM_BYTES, st, ctx->tsdf_map[1], tall_pre, ps_pre,
               ctx->d_xzpart, B, ctx->conv_in);
    } else if (ctx->conv_in_split == 2) {
      using Cf = ConvInCfg<5, 2>;
      launch_k(ctx, conv_in_planes_kernel<5, 2>, dim3(Cf::NT, B, 2), dim3(Cf::THREADS), Cf::SMEM_BYTES, st, ctx->tsdf_map[0], tall_pre, ps_pre,
               ctx->d_xzpart, B, ctx->conv_in);
    } else {
      using Cf = ConvInCfg<5, 1>;
      launch_k(ctx, conv_in_planes_kernel<5, 1>, dim3(Cf::NT, B, 1), dim3(Cf::THREADS), Cf::SMEM_BYTES, st, ctx->tsdf_map[0], tall_pre, ps_pre,
               ctx->d_xzpart, B, ctx->conv_in);
    }
  }
  {
    LaunchScope ls(ctx, "xz_finish", st);
    const int blocks = ceil_div((int)std::max((long)B * 4 * G2, 9 * ctx->flags_stride + 16), 256);
    if (conv_in_ty(B) == 5)
      launch_k(ctx, xz_finish_tall_kernel<G / 5>, dim3(blocks), dim3(256), 0, st, (const float*)ctx->d_xzpart, tall_pre, ps_pre, B, ctx->d_flags,
               (int)(9 * ctx->flags_stride + 16));
    else
      launch_k(ctx, xz_finish_tall_kernel<G>, dim3(blocks), dim3(256), 0, st, (const float*)ctx->d_xzpart, tall_pre, ps_pre, B, ctx->d_flags,
               (int)(9 * ctx->flags_stride + 16));
    ctx->work_slot = 0;
  }
  float *d0c1 = act(ctx, "d0c1"), *d0c2 = act(ctx, "d0c2"), *p0 = act(ctx, "p0"), *d1c1 = act(ctx, "d1c1"),
        *d1c2 = act(ctx, "d1c2"), *p1 = act(ctx, "p1"), *d2c1 = act(ctx, "d2c1"), *d2c2 = act(ctx, "d2c2"),
        *u0 = act(ctx, "u0"), *u0c1 = act(ctx, "u0c1"), *u0c2 = act(ctx, "u0c2"), *u1 = act(ctx, "u1"),
        *u1c1 = act(ctx, "u1c1"), *u1c2 = act(ctx, "u1c2");
  if (ctx->encoder_impl >= 1) {
    unet_tall_forward(ctx, n_img, planes, /*keep_u1c2=*/false, st);
    ctx->last_B = B;
    ctx->last_impl = ctx->encoder_impl;
    CU_TRY(cudaGetLastError());
    return GIGA_OK;
  }
  ctx->last_impl = 0;
  ctx->tr.fwd_valid = false;   // this path overwrites the NCHW activations a pending giga_train_backward would read
  {   // the fp32 FMA-pipe U-Net reads NCHW fp32 planes: expand the pre-split planes (hi + lo * 2^-11 carries 22 significant bits)
    LaunchScope ls(ctx, "tall_to_nchw:pre", st);
    tall_to_nchw_kernel<40, 4><<<ceil_div(n_img * 4 * G2, 256), 256, 0, st>>>(tall_pre, ctx->d_pre, ps_pre, n_img);
  }
  launch_conv<K_d0c1>(ctx, "conv3x3:d0c1", n_img, ctx->d_pre, nullptr, 0, d0c1, nullptr, st);
  launch_conv<K_d0c2>(ctx, "conv3x3:d0c2", n_img, d0c1, nullptr, 1, d0c2, p0, st);
  launch_conv<K_d1c1>(ctx, "conv3x3:d1c1", n_img, p0, nullptr, 2, d1c1, nullptr, st);
  launch_conv<K_d1c2>(ctx, "conv3x3:d1c2", n_img, d1c1, nullptr, 3, d1c2, p1, st);
  launch_conv<K_d2c1>(ctx, "conv3x3:d2c1", n_img, p1, nullptr, 4, d2c1, nullptr, st);
  launch_conv<K_d2c2>(ctx, "conv3x3:d2c2", n_img, d2c1, nullptr, 5, d2c2, nullptr, st);
  launch_convT<K_u0up>(ctx, "convT:u0", n_img, d2c2, 0, u0, st);
  launch_conv<K_u0c1>(ctx, "conv3x3:u0c1", n_img, u0, d1c2, 6, u0c1, nullptr, st);   // cat(from_up, from_down), unet.py:109
  launch_conv<K_u0c2>(ctx, "conv3x3:u0c2", n_img, u0c1, nullptr, 7, u0c2, nullptr, st);
  launch_convT<K_u1up>(ctx, "convT:u1", n_img, u0c2, 1, u1, st);
  launch_conv<K_u1c1>(ctx, "conv3x3:u1c1", n_img, u1, d0c2, 8, u1c1, nullptr, st);
  launch_conv<K_u1c2>(ctx, "conv3x3:u1c2", n_img, u1c1, nullptr, 9, u1c2, nullptr, st);
  {
    LaunchScope ls(ctx, "conv1x1_final", st);
    conv1x1_nhwc_kernel<<<dim3(G2 / F_PIX, n_img), 256, 0, st>>>(u1c2, ctx->d_enc + ctx->el.fin_w, ctx->d_enc + ctx->el.fin_b, planes);
  }
  ctx->last_B = B;
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

}  // extern "C" (re-opened below)

namespace {

// warp-specialised decoder (decoder_ws.cuh): one persistent launch for up to WD_MAX_JOBS jobs
int launch_decode_ws(giga_ctx* ctx, const float* planes, int B, const float* const* pts, const int* Ns, const unsigned* masks, int nsets, bool raw,
                     float* qual, float* rot, float* width, float* occ, cudaStream_t st, const char* name) {
  DecArgs a = {};
  int nj = 0, items = 0;
  auto add = [&](const float* p, int N, int type) {
    DecJob& j = a.job[nj++];
    j.pts = p; j.N = N; j.type = type; j.tiles = ceil_div(N, WD_PTS); j.item0 = items;
    items += B * j.tiles;
  };
  // heavy (3-head) jobs first: items are handed out in index order, so the tail of the launch consists of short items
  for (int s = 0; s < nsets; ++s)
    if ((masks[s] & 7u) == 7u && ctx->d_wblob[0]) add(pts[s], Ns[s], 0);
  for (int s = 0; s < nsets; ++s) {
    const unsigned single = ((masks[s] & 7u) == 7u && ctx->d_wblob[0]) ? (masks[s] & 8u) : (masks[s] & 15u);
    for (int h = 0; h < 4; ++h)
      if (single & (1u << h)) {
        if (nj >= WD_MAX_JOBS) return fail(GIGA_EINVAL, "launch_decode_ws: too many jobs for one launch");
        add(pts[s], Ns[s], 1 + h);
      }
  }
  a.njobs = nj; a.n_items = items; a.B = B; a.raw = raw ? 1u : 0u;
  a.planes = planes; a.hc = ctx->d_hc;
  for (int t = 0; t < 5; ++t) a.wblob[t] = ctx->d_wblob[t];
  a.qual = qual; a.rot = rot; a.width = width; a.occ = occ;
  a.sched = ctx->d_sched;
  const int grid = items < ctx->num_sms ? items : ctx->num_sms;
  a.tl = nullptr;
  a.debug = getenv("GIGA_DEC_DEBUG") ? (unsigned)atoi(getenv("GIGA_DEC_DEBUG")) : 0u;
  if (ctx->timeline_layer && !strcmp(ctx->timeline_layer, "decode")) {
    const size_t n = (size_t)grid * 32;
    if (ctx->d_timeline) cudaFree(ctx->d_timeline);
    cudaMalloc(&ctx->d_timeline, n * 8);
    cudaMemsetAsync(ctx->d_timeline, 0, n * 8, st);
    ctx->timeline_n = (long)n;
    a.tl = ctx->d_timeline;
  }
  LaunchScope ls(ctx, name, st);
  launch_k(ctx, decode_points_ws_kernel, dim3(grid), dim3(WD_THREADS), WD_SMEM_BYTES, st, a);
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

// one decoder launch for up to two jobs (heads at points / heads2 at points2); validation is the callers' job
int launch_decode(giga_ctx* ctx, const float* planes, int B, const float* points, int N, unsigned heads, const float* points2, int N2,
                  unsigned heads2, float* qual, float* rot, float* width, float* occ, cudaStream_t st) {
  const bool two = points2 && N2 > 0;
  if (two && ctx->decoder_impl == 0) return fail(GIGA_EINVAL, "launch_decode: merged jobs need a tensor-core decoder");
  const char* name = two ? "decode_points:grasp+tsdf" : ((heads & 15u) == GIGA_HEAD_TSDF ? "decode_points:tsdf" : "decode_points:grasp");
  if (ctx->decoder_impl == 1) {
    const float* pts[2] = {points, points2};
    const int Ns[2] = {N, N2};
    const unsigned masks[2] = {heads & 15u, heads2 & 15u};
    const bool raw = (heads & 16u) != 0;
    auto njobs = [&](unsigned m) { return ((m & 7u) == 7u && ctx->d_wblob[0]) ? 1 + ((m >> 3) & 1) : __builtin_popcount(m); };
    if (two && njobs(masks[0]) + njobs(masks[1]) > WD_MAX_JOBS) {   // rare head combinations: one launch per point set
      if (int r = launch_decode_ws(ctx, planes, B, pts, Ns, masks, 1, raw, qual, rot, width, occ, st, name)) return r;
      return launch_decode_ws(ctx, planes, B, pts + 1, Ns + 1, masks + 1, 1, raw, qual, rot, width, occ, st, name);
    }
    return launch_decode_ws(ctx, planes, B, pts, Ns, masks, two ? 2 : 1, raw, qual, rot, width, occ, st, name);
  }
  LaunchScope ls(ctx, name, st);
  {
    decode_points_kernel<<<dim3(ceil_div(N, DEC_PTS), B), DEC_PTS, DEC_SMEM_BYTES, st>>>(planes, points, ctx->d_heads, B, N, heads, qual, rot,
                                                                                         width, occ);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

}  // namespace

extern "C" {

int giga_decode(giga_ctx* ctx, const float* planes, int B, const float* points, int N, unsigned heads, float* qual,
                float* rot, float* width, float* occ, void* stream) {
  if (!ctx || !planes || !points || B <= 0 || N <= 0 || !(heads & 15u) || (heads & ~31u))
    return fail(GIGA_EINVAL, "giga_decode: bad argument");
  if (!ctx->committed) return fail(GIGA_ESTATE, "giga_decode: parameters not committed");
  if ((heads & 15u) & ~ctx->heads) return fail(GIGA_ESTATE, "giga_decode: a requested head has no committed parameters");
  if (((heads & GIGA_HEAD_QUAL) && !qual) || ((heads & GIGA_HEAD_ROT) && !rot) || ((heads & GIGA_HEAD_WIDTH) && !width) ||
      ((heads & GIGA_HEAD_TSDF) && !occ))
    return fail(GIGA_EINVAL, "giga_decode: output pointer of a requested head is null");
  if (int r = set_device(ctx)) return r;
  OrderScope order(ctx, (cudaStream_t)stream);
  return launch_decode(ctx, planes, B, points, N, heads, nullptr, 0, 0u, qual, rot, width, occ, (cudaStream_t)stream);
}

int giga_sample_feature(giga_ctx* ctx, const float* planes, int B, const float* points, int N, int mode, float* out,
                        void* stream) {
  if (!ctx || !planes || !points || !out || B <= 0 || N <= 0 || (mode != 0 && mode != 1))
    return fail(GIGA_EINVAL, "giga_sample_feature: bad argument");
  if (int r = set_device(ctx)) return r;
  if (int r = ensure_attrs(ctx)) return r;
  OrderScope order(ctx, (cudaStream_t)stream);
  {
    LaunchScope ls(ctx, "sample_feature", (cudaStream_t)stream);
    sample_feature_kernel<<<dim3(ceil_div(N, DEC_PTS), B), DEC_PTS, SF_SMEM_BYTES, (cudaStream_t)stream>>>(planes, points, B, N, mode, out);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_scene_argmax(giga_ctx* ctx, const float* qual, int B, int N, float* best_val, int* best_idx, void* stream) {
  if (!ctx || !qual || !best_val || !best_idx || B <= 0 || N <= 0) return fail(GIGA_EINVAL, "giga_scene_argmax: bad argument");
  if (int r = set_device(ctx)) return r;
  OrderScope order(ctx, (cudaStream_t)stream);
  {
    LaunchScope ls(ctx, "scene_argmax", (cudaStream_t)stream);
    launch_k(ctx, scene_argmax_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, qual, N, best_val, best_idx);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_forward(giga_ctx* ctx, const float* tsdf, int B, const float* p, int Ng, const float* p_tsdf, int No, float* planes,
                 float* qual, float* rot, float* width, float* occ, float* best_val, int* best_idx, void* stream) {
  if (!ctx || !tsdf || B <= 0) return fail(GIGA_EINVAL, "giga_forward: bad argument");
  const bool grasp = p && Ng > 0, geo = p_tsdf && No > 0;
  if (!grasp && !geo) return fail(GIGA_EINVAL, "giga_forward: no query points");
  if ((best_val || best_idx) && !(grasp && best_val && best_idx && qual))
    return fail(GIGA_EINVAL, "giga_forward: the arg-max needs the grasp heads and both best_val and best_idx");
  if (int r = set_device(ctx)) return r;
  OrderScope order(ctx, (cudaStream_t)stream);
  if (!planes) {
    if (B > ctx->planes_cap) {
      ctx->graph_epoch++;
      CU_TRY(cudaDeviceSynchronize());
      if (ctx->d_planes) cudaFree(ctx->d_planes);
      ctx->d_planes = nullptr;
      ctx->planes_cap = 0;
      CU_TRY(cudaMalloc(&ctx->d_planes, sizeof(float) * 3 * (size_t)B * G2 * C));
      ctx->planes_cap = B;
    }
    planes = ctx->d_planes;
  }
  if (int r = giga_encode(ctx, tsdf, B, planes, stream)) return r;
  const unsigned hm = ctx->heads & (GIGA_HEAD_QUAL | GIGA_HEAD_ROT | GIGA_HEAD_WIDTH);
  if (grasp && !hm) return fail(GIGA_ESTATE, "giga_forward: no grasp head committed (pass p = NULL for giga_geo)");
  if (grasp && geo && ctx->decoder_impl >= 1 && ctx->merge_decode) {   // both point sets in ONE launch (long 3-head tiles first, TSDF tiles fill the tail)
    if (!(ctx->heads & GIGA_HEAD_TSDF)) return fail(GIGA_ESTATE, "giga_forward: no TSDF head committed");
    if (!occ || ((hm & GIGA_HEAD_QUAL) && !qual) || ((hm & GIGA_HEAD_ROT) && !rot) || ((hm & GIGA_HEAD_WIDTH) && !width))
      return fail(GIGA_EINVAL, "giga_forward: output pointer of a requested head is null");
    if (int r = launch_decode(ctx, planes, B, p, Ng, hm, p_tsdf, No, GIGA_HEAD_TSDF, qual, rot, width, occ, (cudaStream_t)stream)) return r;
  } else {
    if (grasp)
      if (int r = giga_decode(ctx, planes, B, p, Ng, hm, qual, rot, width, nullptr, stream)) return r;
    if (geo)
      if (int r = giga_decode(ctx, planes, B, p_tsdf, No, GIGA_HEAD_TSDF, nullptr, nullptr, nullptr, occ, stream)) return r;
  }
  if (best_val)
    if (int r = giga_scene_argmax(ctx, qual, B, Ng, best_val, best_idx, stream)) return r;
  return GIGA_OK;
}

}  // extern "C" (re-opened below)

namespace {

int ensure_slot(giga_ctx* ctx, giga_ctx::HostSlot& s, int B, int Ng, int No) {
  if (B <= s.cap_B && Ng <= s.cap_Ng && No <= s.cap_No) return GIGA_OK;
  CU_TRY(cudaDeviceSynchronize());
  ctx->graph_epoch++;   // captured graphs hold the old staging pointers
  float** ptrs[] = {&s.tsdf, &s.planes, &s.p, &s.pt, &s.qual, &s.rot, &s.width, &s.occ};
  for (float** q : ptrs) { if (*q) cudaFree(*q); *q = nullptr; }
  const int cB = B > s.cap_B ? B : s.cap_B, cg = Ng > s.cap_Ng ? Ng : s.cap_Ng, co = No > s.cap_No ? No : s.cap_No;
  const size_t ng = (size_t)cB * (cg > 0 ? cg : 1), no = (size_t)cB * (co > 0 ? co : 1);
  CU_TRY(cudaMalloc(&s.tsdf, sizeof(float) * (size_t)cB * G3));
  CU_TRY(cudaMalloc(&s.planes, sizeof(float) * 3 * (size_t)cB * G2 * C));
  CU_TRY(cudaMalloc(&s.p, sizeof(float) * ng * 3));
  CU_TRY(cudaMalloc(&s.pt, sizeof(float) * no * 3));
  CU_TRY(cudaMalloc(&s.qual, sizeof(float) * ng));
  CU_TRY(cudaMalloc(&s.rot, sizeof(float) * ng * 4));
  CU_TRY(cudaMalloc(&s.width, sizeof(float) * ng));
  CU_TRY(cudaMalloc(&s.occ, sizeof(float) * no));
  s.cap_B = cB; s.cap_Ng = cg; s.cap_No = co;
  return GIGA_OK;
}

// H2D on `sin`, compute on `sc`, D2H on `sout`; when the three are one stream the event waits are no-ops.
int enqueue_host_request(giga_ctx* ctx, giga_ctx::HostSlot& s, const float* tsdf, int B, const float* p, int Ng, const float* p_tsdf,
                         int No, float* qual, float* rot, float* width, float* occ, cudaStream_t sin, cudaStream_t sc, cudaStream_t sout) {
  const bool grasp = p && Ng > 0, geo = p_tsdf && No > 0;
  CU_TRY(cudaMemcpyAsync(s.tsdf, tsdf, sizeof(float) * (size_t)B * G3, cudaMemcpyHostToDevice, sin));
  if (grasp) CU_TRY(cudaMemcpyAsync(s.p, p, sizeof(float) * (size_t)B * Ng * 3, cudaMemcpyHostToDevice, sin));
  if (geo) CU_TRY(cudaMemcpyAsync(s.pt, p_tsdf, sizeof(float) * (size_t)B * No * 3, cudaMemcpyHostToDevice, sin));
  if (sin != sc) {
    CU_TRY(cudaEventRecord(s.ev_in, sin));
    CU_TRY(cudaStreamWaitEvent(sc, s.ev_in, 0));
  }
  // the kernels: one CUDA-graph launch from the third request of a configuration on (the pipelined path only: sc is then the ctx's own
  // compute stream); the first request runs eagerly (allocations, tensor maps), the second is captured
  bool launched = false;
  if (ctx->use_graph && sc == ctx->st_compute && sc != nullptr && !ctx->timing && !ctx->timeline_layer) {
    const unsigned long long key = 0x9e3779b97f4a7c15ull * (unsigned long long)(B + 1) + 0xc2b2ae3d27d4eb4full * (unsigned long long)(Ng + 1) +
                                   0x165667b19e3779f9ull * (unsigned long long)(No + 1) + ctx->heads;
    if (s.graph && (s.graph_key != key || s.graph_epoch != ctx->graph_epoch)) {
      cudaGraphExecDestroy(s.graph);
      s.graph = nullptr;
      s.seen_key = 0;
    }
    if (!s.graph && s.seen_key == key && s.graph_epoch == ctx->graph_epoch) {   // second request of this configuration: capture
      OrderScope order(ctx, sc);   // orders the capture's replay position; the capture itself runs with the guard off
      const long l0 = ctx->launches;
      CU_TRY(cudaStreamBeginCapture(sc, cudaStreamCaptureModeThreadLocal));
      ctx->capturing = true;
      const int r = giga_forward(ctx, s.tsdf, B, grasp ? s.p : nullptr, Ng, geo ? s.pt : nullptr, No, s.planes, s.qual, s.rot, s.width, s.occ,
                                 nullptr, nullptr, sc);
      ctx->capturing = false;
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(sc, &graph);
      if (r) { if (graph) cudaGraphDestroy(graph); return r; }
      if (ce != cudaSuccess || !graph) return fail(GIGA_ECUDA, std::string("giga_forward_host_submit: graph capture failed: ") + cudaGetErrorString(ce));
      const cudaError_t ie = cudaGraphInstantiate(&s.graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { s.graph = nullptr; return fail(GIGA_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
      s.graph_key = key;
      s.graph_launches = ctx->launches - l0;
      ctx->launches = l0;
    }
    if (s.graph) {
      OrderScope order(ctx, sc);
      CU_TRY(cudaGraphLaunch(s.graph, sc));
      ctx->launches += s.graph_launches;
      launched = true;
    } else {
      s.seen_key = key;
      s.graph_epoch = ctx->graph_epoch;
    }
  }
  if (!launched)
    if (int r = giga_forward(ctx, s.tsdf, B, grasp ? s.p : nullptr, Ng, geo ? s.pt : nullptr, No, s.planes, s.qual, s.rot, s.width, s.occ, nullptr,
                             nullptr, sc))
      return r;
  if (sc != sout) {
    CU_TRY(cudaEventRecord(s.ev_compute, sc));
    CU_TRY(cudaStreamWaitEvent(sout, s.ev_compute, 0));
  }
  if (grasp) {
    CU_TRY(cudaMemcpyAsync(qual, s.qual, sizeof(float) * (size_t)B * Ng, cudaMemcpyDeviceToHost, sout));
    CU_TRY(cudaMemcpyAsync(rot, s.rot, sizeof(float) * (size_t)B * Ng * 4, cudaMemcpyDeviceToHost, sout));
    CU_TRY(cudaMemcpyAsync(width, s.width, sizeof(float) * (size_t)B * Ng, cudaMemcpyDeviceToHost, sout));
  }
  if (geo) CU_TRY(cudaMemcpyAsync(occ, s.occ, sizeof(float) * (size_t)B * No, cudaMemcpyDeviceToHost, sout));
  return GIGA_OK;
}

int check_host_args(const float* tsdf, int B, const float* p, int Ng, const float* p_tsdf, int No, float* qual, float* rot,
                    float* width, float* occ, const char* who) {
  if (!tsdf || B <= 0) return fail(GIGA_EINVAL, std::string(who) + ": bad argument");
  const bool grasp = p && Ng > 0, geo = p_tsdf && No > 0;
  if (!grasp && !geo) return fail(GIGA_EINVAL, std::string(who) + ": no query points");
  if (grasp && (!qual || !rot || !width)) return fail(GIGA_EINVAL, std::string(who) + ": grasp outputs are null");
  if (geo && !occ) return fail(GIGA_EINVAL, std::string(who) + ": occ output is null");
  return GIGA_OK;
}

}  // namespace

extern "C" {

int giga_forward_host(giga_ctx* ctx, const float* tsdf, int B, const float* p, int Ng, const float* p_tsdf, int No,
                      float* qual, float* rot, float* width, float* occ, void* stream) {
  if (!ctx) return fail(GIGA_EINVAL, "giga_forward_host: ctx is null");
  if (int r = check_host_args(tsdf, B, p, Ng, p_tsdf, No, qual, rot, width, occ, "giga_forward_host")) return r;
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  giga_ctx::HostSlot& s = ctx->slot[2];
  if (int r = ensure_slot(ctx, s, B, Ng, No)) return r;
  if (int r = enqueue_host_request(ctx, s, tsdf, B, p, Ng, p_tsdf, No, qual, rot, width, occ, st, st, st)) return r;
  CU_TRY(cudaStreamSynchronize(st));
  return GIGA_OK;
}

int giga_forward_host_submit(giga_ctx* ctx, int slot, const float* tsdf, int B, const float* p, int Ng, const float* p_tsdf, int No,
                             float* qual, float* rot, float* width, float* occ) {
  if (!ctx || slot < 0 || slot > 1) return fail(GIGA_EINVAL, "giga_forward_host_submit: bad ctx / slot (0 or 1)");
  if (int r = check_host_args(tsdf, B, p, Ng, p_tsdf, No, qual, rot, width, occ, "giga_forward_host_submit")) return r;
  if (int r = set_device(ctx)) return r;
  giga_ctx::HostSlot& s = ctx->slot[slot];
  if (s.pending) return fail(GIGA_ESTATE, "giga_forward_host_submit: slot still in flight (call giga_forward_host_wait first)");
  if (!ctx->st_compute) {
    CU_TRY(cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&ctx->st_compute, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking));
  }
  if (!s.ev_in) {
    CU_TRY(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&s.ev_compute, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
  }
  if (int r = ensure_slot(ctx, s, B, Ng, No)) return r;
  if (int r = ensure_workspace(ctx, B)) return r;   // never (re)allocate while the other slot computes
  if (int r = enqueue_host_request(ctx, s, tsdf, B, p, Ng, p_tsdf, No, qual, rot, width, occ, ctx->st_h2d, ctx->st_compute, ctx->st_d2h))
    return r;
  CU_TRY(cudaEventRecord(s.ev_done, ctx->st_d2h));
  s.pending = true;
  return GIGA_OK;
}

int giga_forward_host_wait(giga_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 1) return fail(GIGA_EINVAL, "giga_forward_host_wait: bad ctx / slot");
  giga_ctx::HostSlot& s = ctx->slot[slot];
  if (!s.pending) return fail(GIGA_ESTATE, "giga_forward_host_wait: nothing submitted on this slot");
  CU_TRY(cudaEventSynchronize(s.ev_done));
  s.pending = false;
  return GIGA_OK;
}

// ---- planner post-processing ---------------------------------------------------------------------------
void giga_select_params_default(giga_select_params* p) {
  if (!p) return;
  p->gaussian_sigma = 1.0;
  p->min_width = 0.033f;
  p->max_width = 0.233f;
  p->out_th = 0.5f;
  const double voxel_size = 0.3 / 40;
  p->lim_x = (int)(0.02 / voxel_size);
  p->lim_y = (int)(0.02 / voxel_size);
  p->lim_z = (int)(0.055 / voxel_size);
  p->low_th = 0.5f;
  p->threshold = 0.9f;
  p->force_detection = 0;
  p->max_filter_size = 4;
}

int giga_gaussian_kernel1d(double sigma, int radius, double* out) {
  if (!(sigma > 0) || radius < 0 || radius > PL_MAXR || !out) return fail(GIGA_EINVAL, "giga_gaussian_kernel1d: bad argument");
  const int n = 2 * radius + 1;
  const double sigma2 = sigma * sigma, f = -0.5 / sigma2;
  double a[2 * PL_MAXR + 1];
  for (int i = 0; i < n; ++i) {
    const long x = i - radius;
    a[i] = exp(f * (double)(x * x));
  }
  // numpy's pairwise sum (what `phi_x.sum()` does): plain loop below 8 elements, 8 accumulators from there on
  double sum;
  if (n < 8) {
    sum = 0.0;
    for (int i = 0; i < n; ++i) sum += a[i];
  } else {
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) sum += a[i];
  }
  for (int i = 0; i < n; ++i) out[i] = a[i] / sum;
  return GIGA_OK;
}

}  // extern "C"

namespace {

int make_select(const giga_select_params* prm, SelectParams& S) {
  if (!prm) return fail(GIGA_EINVAL, "giga_select_params is null");
  if (!(prm->gaussian_sigma > 0)) return fail(GIGA_EINVAL, "giga_select_params: gaussian_sigma must be > 0");
  const int radius = (int)(4.0 * prm->gaussian_sigma + 0.5);   // scipy: int(truncate * sd + 0.5), truncate = 4
  if (radius > PL_MAXR) return fail(GIGA_EINVAL, "giga_select_params: gaussian_sigma too large (radius > 8)");
  if (prm->max_filter_size < 1 || prm->max_filter_size > G) return fail(GIGA_EINVAL, "giga_select_params: bad max_filter_size");
  if (prm->lim_x < 0 || prm->lim_y < 0 || prm->lim_z < 0) return fail(GIGA_EINVAL, "giga_select_params: negative bound limit");
  if (int r = giga_gaussian_kernel1d(prm->gaussian_sigma, radius, S.w)) return r;   // symmetric: reversing it (scipy correlates) is a no-op
  S.radius = radius;
  S.min_width = prm->min_width; S.max_width = prm->max_width; S.out_th = prm->out_th;
  S.lim_x = prm->lim_x; S.lim_y = prm->lim_y; S.lim_z = prm->lim_z;
  S.low_th = prm->low_th; S.threshold = prm->threshold;
  S.force_detection = prm->force_detection; S.max_filter_size = prm->max_filter_size;
  return GIGA_OK;
}

int ensure_planner_ws(giga_ctx* ctx, int B) {
  if (B <= ctx->pl_cap_B) return GIGA_OK;
  ctx->graph_epoch++;
  CU_TRY(cudaDeviceSynchronize());
  void** ptrs[] = {(void**)&ctx->d_pl_a, (void**)&ctx->d_pl_b, (void**)&ctx->d_pl_qlow, (void**)&ctx->d_pl_cval, (void**)&ctx->d_pl_cidx,
                   (void**)&ctx->d_pl_flag};
  for (void** q : ptrs) { if (*q) cudaFree(*q); *q = nullptr; }
  ctx->pl_cap_B = 0;
  const size_t vol = sizeof(float) * (size_t)B * G3;
  CU_TRY(cudaMalloc(&ctx->d_pl_a, vol));
  CU_TRY(cudaMalloc(&ctx->d_pl_b, vol));
  CU_TRY(cudaMalloc(&ctx->d_pl_qlow, vol));
  CU_TRY(cudaMalloc(&ctx->d_pl_cval, vol));
  CU_TRY(cudaMalloc(&ctx->d_pl_cidx, vol));
  CU_TRY(cudaMalloc(&ctx->d_pl_flag, sizeof(int) * 2 * (size_t)B));
  ctx->pl_cap_B = B;
  return GIGA_OK;
}

int ensure_detect_ws(giga_ctx* ctx, int B, int K) {
  const size_t out_words_needed = (size_t)B + 7 * (size_t)B * K;
  if (B > ctx->det_cap_B || out_words_needed > ctx->det_out_words) ctx->graph_epoch++;
  if (B > ctx->det_cap_B) {
    CU_TRY(cudaDeviceSynchronize());
    void** ptrs[] = {(void**)&ctx->d_det_pts, (void**)&ctx->d_det_qual, (void**)&ctx->d_det_rot, (void**)&ctx->d_det_width,
                     (void**)&ctx->d_det_tsdf, (void**)&ctx->d_det_tsdfp};
    for (void** q : ptrs) { if (*q) cudaFree(*q); *q = nullptr; }
    ctx->det_cap_B = ctx->det_lat_B = 0;
    const size_t vol = sizeof(float) * (size_t)B * G3;
    CU_TRY(cudaMalloc(&ctx->d_det_pts, vol * 3));
    CU_TRY(cudaMalloc(&ctx->d_det_qual, vol));
    CU_TRY(cudaMalloc(&ctx->d_det_rot, vol * 4));
    CU_TRY(cudaMalloc(&ctx->d_det_width, vol));
    CU_TRY(cudaMalloc(&ctx->d_det_tsdf, vol));
    CU_TRY(cudaMalloc(&ctx->d_det_tsdfp, vol));
    ctx->det_cap_B = B;
  }
  if (out_words_needed > ctx->det_out_words) {   // B + 7*B*K words: compare the word count itself ((B, K) pairs with equal B*K differ in B)
    CU_TRY(cudaDeviceSynchronize());
    if (ctx->d_det_out) cudaFree(ctx->d_det_out);
    ctx->d_det_out = nullptr;
    ctx->det_out_words = 0;
    CU_TRY(cudaMalloc(&ctx->d_det_out, sizeof(int) * out_words_needed));
    ctx->det_out_words = out_words_needed;
    ctx->det_cap_K = B * K;
  }
  return GIGA_OK;
}

}  // namespace

extern "C" {

int giga_select_grasps(giga_ctx* ctx, const float* tsdf, const float* qual, const float* rot, const float* width, int B,
                       const giga_select_params* prm, int K, int* count, float* score, int* index, float* out_rot, float* out_width,
                       float* qual_vol, void* stream) {
  if (!ctx || !tsdf || !qual || !rot || !width || B <= 0 || K <= 0 || !count || !score || !index || !out_rot || !out_width)
    return fail(GIGA_EINVAL, "giga_select_grasps: bad argument");
  if ((reinterpret_cast<uintptr_t>(rot) | reinterpret_cast<uintptr_t>(out_rot)) & 15)
    return fail(GIGA_EINVAL, "giga_select_grasps: rot / out_rot must be 16-byte aligned (quaternions move as float4)");
  SelectParams S;
  if (int r = make_select(prm, S)) return r;
  if (int r = set_device(ctx)) return r;
  if (int r = ensure_planner_ws(ctx, B)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  const int blocks = ceil_div(B * G3, 256);
  int *flag = ctx->d_pl_flag, *cnt = ctx->d_pl_flag + B;
  CU_TRY(cudaMemsetAsync(flag, 0, sizeof(int) * 2 * B, st));
  { LaunchScope ls(ctx, "planner:gauss_x", st); gauss_axis_kernel<0><<<blocks, 256, 0, st>>>(qual, ctx->d_pl_a, B, S); }
  { LaunchScope ls(ctx, "planner:gauss_y", st); gauss_axis_kernel<1><<<blocks, 256, 0, st>>>(ctx->d_pl_a, ctx->d_pl_b, B, S); }
  { LaunchScope ls(ctx, "planner:gauss_z+mask", st);
    gauss_z_mask_kernel<<<blocks, 256, 0, st>>>(ctx->d_pl_b, tsdf, width, qual_vol, ctx->d_pl_qlow, flag, B, S); }
  { LaunchScope ls(ctx, "planner:nms", st);
    grasp_nms_kernel<<<blocks, 256, 0, st>>>(ctx->d_pl_qlow, flag, cnt, ctx->d_pl_cval, ctx->d_pl_cidx, B, S); }
  { LaunchScope ls(ctx, "planner:rank", st);
    grasp_rank_kernel<<<dim3(G3 / 256, B), 256, 0, st>>>(cnt, flag, ctx->d_pl_cval, ctx->d_pl_cidx, rot, width, K, count, score, index,
                                                        out_rot, out_width, S); }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_ctx_set_lattice(giga_ctx* ctx, const float* pos, int N) {
  if (!ctx || !pos || N != G3) return fail(GIGA_EINVAL, "giga_ctx_set_lattice: the planner lattice is [64000][3] (40^3 volumes)");
  if (int r = set_device(ctx)) return r;
  CU_TRY(cudaDeviceSynchronize());
  if (!ctx->d_lattice) CU_TRY(cudaMalloc(&ctx->d_lattice, sizeof(float) * 3 * G3));
  CU_TRY(cudaMemcpy(ctx->d_lattice, pos, sizeof(float) * 3 * G3, cudaMemcpyHostToDevice));
  ctx->det_lat_B = 0;   // re-broadcast on the next giga_detect
  ctx->graph_epoch++;
  return GIGA_OK;
}

int giga_detect(giga_ctx* ctx, const float* tsdf, const float* tsdf_process, int B, const giga_select_params* prm, int K, int* count,
                float* score, int* index, float* out_rot, float* out_width, void* stream) {
  if (!ctx || !tsdf || B <= 0 || K <= 0) return fail(GIGA_EINVAL, "giga_detect: bad argument");
  if (!ctx->d_lattice) return fail(GIGA_ESTATE, "giga_detect: no query lattice (call giga_ctx_set_lattice first)");
  if (int r = set_device(ctx)) return r;
  if (int r = ensure_detect_ws(ctx, B, K)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  if (ctx->det_lat_B < B) {
    LaunchScope ls(ctx, "planner:broadcast_lattice", st);
    broadcast_points_kernel<<<ceil_div(3 * G3, 256), 256, 0, st>>>(ctx->d_lattice, ctx->d_det_pts, 3 * G3, ctx->det_cap_B);
    ctx->det_lat_B = ctx->det_cap_B;
  }
  if (int r = giga_forward(ctx, tsdf, B, ctx->d_det_pts, G3, nullptr, 0, nullptr, ctx->d_det_qual, ctx->d_det_rot, ctx->d_det_width, nullptr,
                           nullptr, nullptr, stream))
    return r;
  return giga_select_grasps(ctx, tsdf_process ? tsdf_process : tsdf, ctx->d_det_qual, ctx->d_det_rot, ctx->d_det_width, B, prm, K, count,
                            score, index, out_rot, out_width, nullptr, stream);
}

int giga_detect_host(giga_ctx* ctx, const float* tsdf, const float* tsdf_process, int B, const giga_select_params* prm, int K, int* count,
                     float* score, int* index, float* out_rot, float* out_width, void* stream) {
  if (!ctx || !tsdf || !prm || B <= 0 || K <= 0 || !count || !score || !index || !out_rot || !out_width)
    return fail(GIGA_EINVAL, "giga_detect_host: bad argument");
  if (int r = set_device(ctx)) return r;
  if (int r = ensure_detect_ws(ctx, B, K)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);   // covers the graph replay on `st`; the capture itself runs with the guard off
  const size_t vol = sizeof(float) * (size_t)B * G3, bk = (size_t)B * K, out_words = 7 * bk + B;
  // pinned staging: the call is ONE H2D (+1 with a separate tsdf_process), the kernels, ONE D2H of B*(1+7K) words
  if (B <= 4 && 2 * vol > ctx->h_det_in_cap) {
    CU_TRY(cudaDeviceSynchronize());
    if (ctx->h_det_in) cudaFreeHost(ctx->h_det_in);
    ctx->h_det_in = nullptr; ctx->h_det_in_cap = 0;
    CU_TRY(cudaMallocHost(&ctx->h_det_in, 2 * vol));
    ctx->h_det_in_cap = 2 * vol;
    ctx->graph_epoch++;
  }
  if (out_words * 4 > ctx->h_det_out_cap) {
    CU_TRY(cudaDeviceSynchronize());
    if (ctx->h_det_out) cudaFreeHost(ctx->h_det_out);
    ctx->h_det_out = nullptr; ctx->h_det_out_cap = 0;
    CU_TRY(cudaMallocHost(&ctx->h_det_out, out_words * 4));
    ctx->h_det_out_cap = out_words * 4;
    ctx->graph_epoch++;
  }
  // small batches (the latency regime) stage the input through the pinned buffer so that the whole call can be a graph;
  // large batches copy straight from the caller's memory (an extra 8 MB host memcpy would cost more than the launches)
  const bool stage_in = B <= 4;
  const void* src = tsdf;
  const void* src_p = tsdf_process;
  if (stage_in) {
    memcpy(ctx->h_det_in, tsdf, vol);
    if (tsdf_process) memcpy(reinterpret_cast<char*>(ctx->h_det_in) + vol, tsdf_process, vol);
    src = ctx->h_det_in;
    src_p = reinterpret_cast<char*>(ctx->h_det_in) + vol;
  }
  float* d_rot = reinterpret_cast<float*>(ctx->d_det_out);   // [B][K][4] rot | score | width | index | count (rot first: float4 stores)
  float* d_score = d_rot + 4 * bk;
  float* d_width = d_score + bk;
  int* d_index = reinterpret_cast<int*>(d_width + bk);
  int* d_count = d_index + bk;
  auto enqueue = [&](cudaStream_t s) -> int {
    CU_TRY(cudaMemcpyAsync(ctx->d_det_tsdf, src, vol, cudaMemcpyHostToDevice, s));
    if (tsdf_process) CU_TRY(cudaMemcpyAsync(ctx->d_det_tsdfp, src_p, vol, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemsetAsync(ctx->d_det_out, 0, out_words * 4, s));   // entries past count[b] read as zero, not as the previous call's grasps
    if (int r = giga_detect(ctx, ctx->d_det_tsdf, tsdf_process ? ctx->d_det_tsdfp : nullptr, B, prm, K, d_count, d_score, d_index, d_rot, d_width, s))
      return r;
    CU_TRY(cudaMemcpyAsync(ctx->h_det_out, ctx->d_det_out, out_words * 4, cudaMemcpyDeviceToHost, s));
    return GIGA_OK;
  };
  // CUDA-graph replay: the first call of a configuration runs eagerly (allocations, lattice broadcast), the second
  // captures the same enqueue sequence, later ones replay it with one cudaGraphLaunch.
  bool done = false;
  if (ctx->use_graph && stage_in && !ctx->timing && !ctx->timeline_layer) {
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { for (size_t i = 0; i < n; ++i) key = (key ^ reinterpret_cast<const unsigned char*>(p)[i]) * 1099511628211ull; };
    const int cfg[4] = {B, K, tsdf_process ? 1 : 0, (int)ctx->heads};
    mix(cfg, sizeof cfg);
    mix(prm, sizeof *prm);
    giga_ctx::DetGraph* g = nullptr;
    for (auto& e : ctx->det_graphs)
      if (e.key == key) g = &e;
    if (g && g->epoch != ctx->graph_epoch) {   // stale: pointers or baked parameters changed
      if (g->exec) cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
      g->epoch = ctx->graph_epoch;
      g = nullptr;                              // this call runs eagerly again
    } else if (!g) {
      if (ctx->det_graphs.size() >= 16) {
        for (auto& e : ctx->det_graphs)
          if (e.exec) cudaGraphExecDestroy(e.exec);
        ctx->det_graphs.clear();
      }
      ctx->det_graphs.push_back({key, ctx->graph_epoch + 1, nullptr, 0});   // epoch + 1: never matches -> next call re-registers after the eager run
      ctx->det_graphs.back().epoch = ~0ull;
      g = nullptr;
    }
    if (g && !g->exec) {
      if (!ctx->st_graph) CU_TRY(cudaStreamCreateWithFlags(&ctx->st_graph, cudaStreamNonBlocking));
      const long l0 = ctx->launches;
      CU_TRY(cudaStreamBeginCapture(ctx->st_graph, cudaStreamCaptureModeThreadLocal));
      ctx->capturing = true;
      const int r = enqueue(ctx->st_graph);
      ctx->capturing = false;
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(ctx->st_graph, &graph);
      if (r) { if (graph) cudaGraphDestroy(graph); return r; }
      if (ce != cudaSuccess || !graph) return fail(GIGA_ECUDA, std::string("giga_detect_host: graph capture failed: ") + cudaGetErrorString(ce));
      const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { g->exec = nullptr; return fail(GIGA_ECUDA, std::string("giga_detect_host: cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
      g->launches = ctx->launches - l0;
      ctx->launches = l0;                       // capture enqueued nothing; replays are counted below
    }
    if (g && g->exec) {
      CU_TRY(cudaGraphLaunch(g->exec, st));
      ctx->launches += g->launches;
      done = true;
    }
  }
  if (!done) {
    const unsigned long long e0 = ctx->graph_epoch;
    if (int r = enqueue(st)) return r;
    if (ctx->use_graph)                          // register the configuration: the epoch after this eager run is the valid one
      for (auto& e : ctx->det_graphs)
        if (e.epoch == ~0ull) e.epoch = ctx->graph_epoch;
    (void)e0;
  }
  CU_TRY(cudaStreamSynchronize(st));
  const int* ho = ctx->h_det_out;
  memcpy(out_rot, ho, sizeof(float) * 4 * bk);
  memcpy(score, ho + 4 * bk, sizeof(float) * bk);
  memcpy(out_width, ho + 5 * bk, sizeof(float) * bk);
  memcpy(index, ho + 6 * bk, sizeof(int) * bk);
  memcpy(count, ho + 7 * bk, sizeof(int) * B);
  return GIGA_OK;
}

int giga_vgn_forward(giga_ctx* ctx, const float* tsdf, int B, float* qual, float* rot, float* width, void* stream) {
  if (!ctx || !tsdf || !qual || !rot || !width || B <= 0) return fail(GIGA_EINVAL, "giga_vgn_forward: bad argument");
  if (!ctx->committed || !ctx->has_vgn) return fail(GIGA_ESTATE, "giga_vgn_forward: VGN parameters not committed");
  if (reinterpret_cast<uintptr_t>(rot) & 15) return fail(GIGA_EINVAL, "giga_vgn_forward: rot must be 16-byte aligned");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  long act_per = 0;
  for (long a : kVgnAct) act_per += a;
  if (B > ctx->vgn_cap_B) {
    CU_TRY(cudaDeviceSynchronize());
    if (ctx->d_vgn_act) cudaFree(ctx->d_vgn_act);
    ctx->d_vgn_act = nullptr;
    ctx->vgn_cap_B = 0;
    CU_TRY(cudaMalloc(&ctx->d_vgn_act, sizeof(float) * act_per * B));
    ctx->vgn_cap_B = B;
    ctx->graph_epoch++;
  }
  const VgnLayout VL = make_vgn_layout();
  const float* P = ctx->d_vgn;
  float* a[6];
  {
    float* q = ctx->d_vgn_act;
    for (int i = 0; i < 6; ++i) { a[i] = q; q += kVgnAct[i] * B; }
  }
  auto grid = [&](int v) { return dim3((unsigned)ceil_div(v, 128), (unsigned)B); };
  // <CIN, COUT, K, STRIDE, UPIN, DIN, DOUT, CI_CH, EPI>
  { LaunchScope ls(ctx, "vgn:enc1", st); conv3d_kernel<1, 16, 5, 2, false, 40, 20, 1, 0><<<grid(8000), 128, 0, st>>>(tsdf, P + VL.w[0], P + VL.b[0], a[0], nullptr, nullptr); }
  { LaunchScope ls(ctx, "vgn:enc2", st); conv3d_kernel<16, 32, 3, 2, false, 20, 10, 8, 0><<<grid(1000), 128, 0, st>>>(a[0], P + VL.w[1], P + VL.b[1], a[1], nullptr, nullptr); }
  { LaunchScope ls(ctx, "vgn:enc3", st); conv3d_kernel<32, 64, 3, 2, false, 10, 5, 4, 0><<<grid(125), 128, 0, st>>>(a[1], P + VL.w[2], P + VL.b[2], a[2], nullptr, nullptr); }
  { LaunchScope ls(ctx, "vgn:dec1", st); conv3d_kernel<64, 64, 3, 1, false, 5, 5, 4, 0><<<grid(125), 128, 0, st>>>(a[2], P + VL.w[3], P + VL.b[3], a[3], nullptr, nullptr); }
  { LaunchScope ls(ctx, "vgn:dec2", st); conv3d_kernel<64, 32, 3, 1, true, 5, 10, 8, 0><<<grid(1000), 128, 0, st>>>(a[3], P + VL.w[4], P + VL.b[4], a[4], nullptr, nullptr); }
  { LaunchScope ls(ctx, "vgn:dec3", st); conv3d_kernel<32, 16, 5, 1, true, 10, 20, 4, 0><<<grid(8000), 128, 0, st>>>(a[4], P + VL.w[5], P + VL.b[5], a[5], nullptr, nullptr); }
  { LaunchScope ls(ctx, "vgn:heads", st); conv3d_kernel<16, 8, 5, 1, true, 20, 40, 8, 1><<<grid(G3), 128, 0, st>>>(a[5], P + VL.w[6], P + VL.b[6], qual, rot, width); }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_loss(giga_ctx* ctx, const float* label_pred, const float* rot_pred, const float* width_pred, const float* occ_pred, const float* label,
              const float* rotations, const float* width, const float* occ, int B, int M, float* loss_out, float* g_label, float* g_rot, float* g_width,
              float* g_occ, void* stream) {
  if (!ctx || !label_pred || !rot_pred || !width_pred || !label || !rotations || !width || !loss_out || B <= 0 || M < 0 || (M > 0 && (!occ_pred || !occ)))
    return fail(GIGA_EINVAL, "giga_loss: bad argument");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  if (B > ctx->loss_cap_B) {
    CU_TRY(cudaDeviceSynchronize());
    if (ctx->d_loss_part) cudaFree(ctx->d_loss_part);
    ctx->d_loss_part = nullptr;
    ctx->loss_cap_B = 0;
    CU_TRY(cudaMalloc(&ctx->d_loss_part, sizeof(float) * 4 * B));
    if (!ctx->d_loss_done) {
      CU_TRY(cudaMalloc(&ctx->d_loss_done, sizeof(unsigned)));
      CU_TRY(cudaMemset(ctx->d_loss_done, 0, sizeof(unsigned)));
    }
    ctx->loss_cap_B = B;
  }
  {
    LaunchScope ls(ctx, "train:loss", st);
    giga_loss_kernel<<<B, 256, 0, st>>>(label_pred, rot_pred, width_pred, occ_pred, label, rotations, width, occ, B, M, ctx->d_loss_part, ctx->d_loss_done,
                                        loss_out, g_label, g_rot, g_width, g_occ);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_adam_step(giga_ctx* ctx, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int step, double lr, double beta1,
                   double beta2, double eps, double weight_decay, void* stream) {
  if (!ctx || !param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return fail(GIGA_EINVAL, "giga_adam_step: bad argument");
  if (!(lr >= 0) || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1) || !(eps >= 0) || !(weight_decay >= 0))
    return fail(GIGA_EINVAL, "giga_adam_step: invalid hyper-parameter");     // torch/optim/adam.py:52-62 raises ValueError for the same
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(GIGA_EINVAL, "giga_adam_step: buffers must be 16-byte aligned");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  // the scalar arithmetic of torch/optim/adam.py _single_tensor_adam, in Python floats (double), rounded to fp32 where ATen does
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  AdamScalars S;
  S.beta1 = (float)beta1; S.beta2 = (float)beta2; S.one_minus_beta1 = (float)(1.0 - beta1); S.one_minus_beta2 = (float)(1.0 - beta2);
  S.step_size = (float)(lr / bc1); S.bc2_sqrt = (float)std::sqrt(bc2); S.eps = (float)eps; S.weight_decay = (float)weight_decay;
  const long n4 = n >> 2;
  const int blocks = (int)std::max(1L, std::min((n4 + 255) / 256, (long)ctx->num_sms * 8));
  {
    LaunchScope ls(ctx, "train:adam", st);
    adam_step_kernel<<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, S);
  }
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

int giga_mise_sweep(giga_ctx* ctx, const float* planes, int resolution0, int upsampling_steps, double threshold, double box_size, float* value_grid,
                    int* stats, void* stream) {
  if (!ctx || !planes || !value_grid || resolution0 < 1 || upsampling_steps < 1 || upsampling_steps > 6)
    return fail(GIGA_EINVAL, "giga_mise_sweep: bad argument (upsampling_steps 1..6; use giga_decode on a regular grid for 0)");
  const long R = (long)resolution0 << upsampling_steps;
  if (R > 512) return fail(GIGA_EINVAL, "giga_mise_sweep: resolution0 * 2^upsampling_steps must be <= 512");
  if (!ctx->committed || !(ctx->heads & GIGA_HEAD_TSDF)) return fail(GIGA_ESTATE, "giga_mise_sweep: no committed TSDF head (decoder_tsdf)");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  const long L = R + 1, np = L * L * L, nc = R * R * R;
  if (ctx->mise_R != (int)R) {
    CU_TRY(cudaDeviceSynchronize());
    void** mp[] = {(void**)&ctx->d_mise_pstate, (void**)&ctx->d_mise_level, (void**)&ctx->d_mise_active, (void**)&ctx->d_mise_val, (void**)&ctx->d_mise_qpts,
                   (void**)&ctx->d_mise_occ, (void**)&ctx->d_mise_qidx, (void**)&ctx->d_mise_count};
    for (void** q : mp) { if (*q) cudaFree(*q); *q = nullptr; }
    ctx->mise_R = 0;
    CU_TRY(cudaMalloc(&ctx->d_mise_pstate, np));
    CU_TRY(cudaMalloc(&ctx->d_mise_level, nc));
    CU_TRY(cudaMalloc(&ctx->d_mise_active, nc));
    CU_TRY(cudaMalloc(&ctx->d_mise_val, sizeof(float) * np));
    CU_TRY(cudaMalloc(&ctx->d_mise_qpts, sizeof(float) * 3 * np));
    CU_TRY(cudaMalloc(&ctx->d_mise_occ, sizeof(float) * np));
    CU_TRY(cudaMalloc(&ctx->d_mise_qidx, sizeof(int) * np));
    CU_TRY(cudaMalloc(&ctx->d_mise_count, sizeof(int)));
    ctx->mise_R = (int)R;
  }
  const MiseDims d = {(int)R, (int)L, upsampling_steps};
  const int gp = (int)((np + 255) / 256), gc = (int)((nc + 255) / 256);
  { LaunchScope ls(ctx, "mise:init", st);
    mise_init_kernel<<<gp, 256, 0, st>>>(d, ctx->d_mise_pstate, ctx->d_mise_val, ctx->d_mise_level, ctx->d_mise_active, ctx->d_mise_count); }
  int iters = 0;
  long total = 0;
  for (;; ++iters) {
    if (iters > 64) return fail(GIGA_ESTATE, "giga_mise_sweep: no convergence after 64 iterations");
    { LaunchScope ls(ctx, "mise:collect", st);
      mise_collect_kernel<<<gp, 256, 0, st>>>(d, ctx->d_mise_pstate, box_size, ctx->d_mise_count, ctx->d_mise_qidx, ctx->d_mise_qpts, (int)np); }
    int n = 0;   // the one host round trip per level: the size of the next decoder launch (the reference moves every value through the host)
    CU_TRY(cudaMemcpyAsync(&n, ctx->d_mise_count, sizeof n, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (n == 0) break;
    total += n;
    if (int r = launch_decode(ctx, planes, 1, ctx->d_mise_qpts, n, GIGA_HEAD_TSDF, nullptr, 0, 0u, nullptr, nullptr, nullptr, ctx->d_mise_occ, st)) return r;
    { LaunchScope ls(ctx, "mise:scatter", st);
      mise_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_mise_qidx, ctx->d_mise_occ, n, ctx->d_mise_pstate, ctx->d_mise_val, ctx->d_mise_count); }
    { LaunchScope ls(ctx, "mise:mark", st);
      mise_mark_kernel<<<gc, 256, 0, st>>>(d, ctx->d_mise_pstate, ctx->d_mise_val, ctx->d_mise_level, ctx->d_mise_active, threshold); }
    { LaunchScope ls(ctx, "mise:subdivide", st);
      mise_apply_kernel<<<gc, 256, 0, st>>>(d, ctx->d_mise_pstate, ctx->d_mise_level, ctx->d_mise_active); }
  }
  const int gl = (int)((L * L + 255) / 256);
  { LaunchScope ls(ctx, "mise:dense", st);
    mise_dense_kernel<0><<<gl, 256, 0, st>>>(d, ctx->d_mise_pstate, ctx->d_mise_val, value_grid);
    mise_dense_kernel<1><<<gl, 256, 0, st>>>(d, ctx->d_mise_pstate, ctx->d_mise_val, value_grid);
    mise_dense_kernel<2><<<gl, 256, 0, st>>>(d, ctx->d_mise_pstate, ctx->d_mise_val, value_grid);
    ctx->launches += 2; }
  CU_TRY(cudaGetLastError());
  if (stats) { stats[0] = iters; stats[1] = (int)total; }
  return GIGA_OK;
}

long giga_ctx_launch_count(const giga_ctx* ctx) { return ctx ? ctx->launches : 0; }

long giga_ctx_overflow_count(giga_ctx* ctx, int reset) {
  if (!ctx) return fail(GIGA_EINVAL, "giga_ctx_overflow_count: ctx is null");
  if (!ctx->d_sched) return 0;
  if (int r = set_device(ctx)) return r;
  CU_TRY(cudaDeviceSynchronize());
  unsigned n = 0;
  CU_TRY(cudaMemcpy(&n, ctx->d_sched + 2, sizeof n, cudaMemcpyDeviceToHost));
  if (reset && n) CU_TRY(cudaMemset(ctx->d_sched + 2, 0, sizeof n));
  if (n) g_err = "giga: " + std::to_string(n) + " decoder outputs were not finite: an activation or plane feature left fp16's +-65504 operand range "
                 "(or an input was not finite); use decoder_impl 0 (fp32 FMA pipe) for such weights";
  return (long)n;
}

int giga_ctx_set_option(giga_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return fail(GIGA_EINVAL, "giga_ctx_set_option: bad argument");
  ctx->graph_epoch++;
  if (!strcmp(key, "graph")) {
    if (value != 0 && value != 1) return fail(GIGA_EINVAL, "graph must be 0 or 1");
    ctx->use_graph = value;
    return GIGA_OK;
  }
  if (!strcmp(key, "decoder_impl")) {
    if (value != 0 && value != 1) return fail(GIGA_EINVAL, "decoder_impl must be 0 (fp32 FMA pipe) or 1 (warp-specialised tcgen05 3xFP16)");
    ctx->decoder_impl = value;
    return GIGA_OK;
  }
  if (!strcmp(key, "dynamic_items")) {
    if (value != 0 && value != 1) return fail(GIGA_EINVAL, "dynamic_items must be 0 or 1");
    ctx->dynamic_items = value;
    return GIGA_OK;
  }
  if (!strcmp(key, "tile_deps")) {
    if (value != 0 && value != 1) return fail(GIGA_EINVAL, "tile_deps must be 0 or 1");
    ctx->tile_deps = value;
    return GIGA_OK;
  }
  if (!strcmp(key, "pdl")) {
    if (value != 0 && value != 1) return fail(GIGA_EINVAL, "pdl must be 0 or 1");
    ctx->pdl = value;
    return GIGA_OK;
  }
  if (!strcmp(key, "train_forward_impl")) {
    if (value < 0 || value > 1) return fail(GIGA_EINVAL, "train_forward_impl must be 0 (fp32 FMA pipe) or 1 (tcgen05 kernels on device-packed operands)");
    ctx->tr.fwd_impl = value;
    return GIGA_OK;
  }
  if (!strcmp(key, "encoder_impl")) {
    if (value < 0 || value > 1) return fail(GIGA_EINVAL, "encoder_impl must be 0 (fp32 FMA pipe) or 1 (tcgen05 3xFP16, persistent)");
    ctx->encoder_impl = value;
    return GIGA_OK;
  }
  return fail(GIGA_EINVAL, std::string("giga_ctx_set_option: unknown key '") + key + "'");
}

int giga_ctx_set_timing(giga_ctx* ctx, int enabled) {
  if (!ctx) return fail(GIGA_EINVAL, "giga_ctx_set_timing: ctx is null");
  ctx->timing = enabled != 0;
  return GIGA_OK;
}

long giga_ctx_timing_report(giga_ctx* ctx, char* buf, long cap) {
  if (!ctx || !buf || cap <= 0) return fail(GIGA_EINVAL, "giga_ctx_timing_report: bad argument");
  if (int r = set_device(ctx)) return r;
  std::map<std::string, std::pair<long, double>> agg;
  std::vector<std::string> order;
  for (auto& t : ctx->timed) {
    CU_TRY(cudaEventSynchronize(t.b));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, t.a, t.b));
    if (!agg.count(t.name)) order.push_back(t.name);
    auto& e = agg[t.name];
    e.first++;
    e.second += ms;
    ctx->ev_pool.push_back(t.a);
    ctx->ev_pool.push_back(t.b);
  }
  ctx->timed.clear();
  std::string out;
  for (auto& n : order) {
    char line[160];
    snprintf(line, sizeof line, "%s %ld %.6f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  if ((long)out.size() + 1 > cap) return fail(GIGA_EINVAL, "giga_ctx_timing_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return (long)out.size();
}

long giga_debug_copy(giga_ctx* ctx, const char* name, float* dst, long capacity, void* stream) {
  if (!ctx || !name || !dst) return fail(GIGA_EINVAL, "giga_debug_copy: bad argument");
  if (ctx->last_B <= 0 && strcmp(name, "timeline")) return fail(GIGA_ESTATE, "giga_debug_copy: no giga_encode call yet");
  if (int r = set_device(ctx)) return r;
  OrderScope order(ctx, (cudaStream_t)stream);
  const float* src = nullptr;
  long numel = 0;
  if (!strcmp(name, "layer_times")) {
    if (!ctx->d_layer_times) return fail(GIGA_ESTATE, "giga_debug_copy: set GIGA_LAYER_TIMES=1");
    if (128 > capacity) return fail(GIGA_EINVAL, "giga_debug_copy: destination too small");
    CU_TRY(cudaMemcpyAsync(dst, ctx->d_layer_times, 64 * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 128;
  }
  if (!strcmp(name, "timeline")) {   // u64 stamps reinterpreted as pairs of floats (debug only)
    if (!ctx->d_timeline) return fail(GIGA_ESTATE, "giga_debug_copy: no timeline recorded (set GIGA_TIMELINE=<kernel name>)");
    if (ctx->timeline_n * 2 > capacity) return fail(GIGA_EINVAL, "giga_debug_copy: destination too small");
    CU_TRY(cudaMemcpyAsync(dst, ctx->d_timeline, ctx->timeline_n * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ctx->timeline_n * 2;
  }
  if (!strcmp(name, "pre")) {   // the planes before the U-Net exist only as TALL pre-split operands: expand into the NCHW scratch
    src = ctx->d_pre;
    numel = 3L * ctx->last_B * C * G2;
    const int n_img = 3 * ctx->last_B;
    tall_to_nchw_kernel<40, 4><<<ceil_div(n_img * 4 * G2, 256), 256, 0, (cudaStream_t)stream>>>(ctx->d_tall[0], ctx->d_pre, ctx->tall_ps[0], n_img);
  } else {
    for (int i = 0; i < kNumActs; ++i)
      if (!strcmp(kActs[i].name, name)) {
        src = ctx->d_act[i];
        numel = 3L * ctx->last_B * kActs[i].ch * kActs[i].hw * kActs[i].hw;
      }
  }
  if (!src) return fail(GIGA_EINVAL, std::string("giga_debug_copy: unknown buffer '") + name + "'");
  if (ctx->last_impl >= 1 && strcmp(name, "pre")) {
    if (!strcmp(name, "u1c2"))
      return fail(GIGA_ESTATE, "giga_debug_copy: 'u1c2' is fused away by the tensor-core encoder (encoder_impl=1)");
    // the tensor-core encoder keeps activations in the TALL pre-split layout: convert into the NCHW scratch first
    int idx = -1;
    for (int i = 0; i < kNumActs; ++i)
      if (!strcmp(kActs[i].name, name)) idx = i;
    const int n_img = 3 * ctx->last_B, hw = kActs[idx].hw, c8 = kActs[idx].ch / 8;
    const float* tsrc = ctx->d_tall[1 + idx];
    const long ps = ctx->tall_ps[1 + idx];
    float* scratch = ctx->d_act[idx];
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = ceil_div(n_img * c8 * hw * hw, 256);
    if (hw == 40 && c8 == 4) tall_to_nchw_kernel<40, 4><<<blocks, 256, 0, st>>>(tsrc, scratch, ps, n_img);
    else if (hw == 20 && c8 == 4) tall_to_nchw_kernel<20, 4><<<blocks, 256, 0, st>>>(tsrc, scratch, ps, n_img);
    else if (hw == 20 && c8 == 8) tall_to_nchw_kernel<20, 8><<<blocks, 256, 0, st>>>(tsrc, scratch, ps, n_img);
    else if (hw == 10 && c8 == 8) tall_to_nchw_kernel<10, 8><<<blocks, 256, 0, st>>>(tsrc, scratch, ps, n_img);
    else if (hw == 10 && c8 == 16) tall_to_nchw_kernel<10, 16><<<blocks, 256, 0, st>>>(tsrc, scratch, ps, n_img);
    else return fail(GIGA_EINVAL, "giga_debug_copy: unexpected activation shape");
  }
  if (numel > capacity) return fail(GIGA_EINVAL, "giga_debug_copy: destination too small");
  CU_TRY(cudaMemcpyAsync(dst, src, sizeof(float) * numel, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return numel;
}

}  // extern "C"

#include "train_api.cuh"
#include "tsdf_api.cuh"
