// C-ABI entry points of the native training step (giga_train_bind / giga_train_forward / giga_train_backward); kernels in train_bwd.cuh.
// Included at the end of giga_api.cu (same translation unit: the kernels' __constant__ symbol and the ctx helpers live there).
#pragma once

namespace {

// slot table: [0] conv_in.weight [1] conv_in.bias; 2 + 2i / 3 + 2i: U-Net conv i weight / bias (kConvName order); 22 + 2i / 23 + 2i: upconv i;
// 26 / 27 conv_final; 28 + 34 h: head h = fc_p (w, b), 5 x (fc_c, fc_0, fc_1) (w, b), fc_out (w, b)
constexpr int TS_HEAD0 = 28, TS_HEAD = 34;

std::string train_slot_name(int s) {
  if (s == 0) return "encoder.conv_in.weight";
  if (s == 1) return "encoder.conv_in.bias";
  if (s < 22) return std::string("encoder.unet.") + kConvName[(s - 2) / 2] + ((s & 1) ? ".bias" : ".weight");
  if (s < 26) return "encoder.unet.up_convs." + std::to_string((s - 22) / 2) + ".upconv" + ((s & 1) ? ".bias" : ".weight");
  if (s < 28) return std::string("encoder.unet.conv_final") + ((s & 1) ? ".bias" : ".weight");
  const int h = (s - TS_HEAD0) / TS_HEAD, r = (s - TS_HEAD0) % TS_HEAD;
  const std::string pre = std::string("decoder_") + kHeadName[h] + ".";
  const char* wb = (r & 1) ? ".bias" : ".weight";
  if (r < 2) return pre + "fc_p" + wb;
  if (r >= 32) return pre + "fc_out" + wb;
  const int i = (r - 2) / 6, f = ((r - 2) % 6) / 2;
  if (f == 0) return pre + "fc_c." + std::to_string(i) + wb;
  return pre + "blocks." + std::to_string(i) + ".fc_" + std::to_string(f - 1) + wb;
}

// data-gradient instances of the forward conv kernel (HW, C of the incoming gradient, 0, C of the outgoing gradient, ...)
using G_40 = Conv3x3Cfg<40, 32, 0, 32, 4, 32, 4, 16, false, 1>;
using G_20a = Conv3x3Cfg<20, 64, 0, 32, 4, 32, 4, 16, false, 1>;
using G_20b = Conv3x3Cfg<20, 64, 0, 64, 4, 64, 4, 16, false, 1>;
using G_10a = Conv3x3Cfg<10, 128, 0, 64, 10, 32, 4, 16, false, 1>;
using G_10b = Conv3x3Cfg<10, 128, 0, 128, 10, 32, 4, 16, false, 1>;
using W_40 = WgradCfg<40, 4>;
using W_20 = WgradCfg<20, 4>;
using W_10 = WgradCfg<10, 10>;
using TB_u0 = ConvTBwdCfg<10, 128, 64>;
using TB_u1 = ConvTBwdCfg<20, 64, 32>;

template <class K> constexpr int convT_wgrad_smem() { return (32 * K::PSX + 32 * 2 * K::R * K::GSW) * 4; }
template <class K> constexpr int convT_dgrad_smem() { return (32 * K::COUT * 4 + 32 * K::PSG) * 4; }

int train_attrs(giga_ctx* ctx) {
  static bool done[64] = {};
  if (done[ctx->device & 63]) return GIGA_OK;
#define SET_G(K) CU_TRY(cudaFuncSetAttribute(conv3x3_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES))
  SET_G(G_40); SET_G(G_20a); SET_G(G_20b); SET_G(G_10a); SET_G(G_10b);
#undef SET_G
  CU_TRY(cudaFuncSetAttribute(conv3x3_wgrad_kernel<W_40>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_40::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(conv3x3_wgrad_kernel<W_20>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_20::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(conv3x3_wgrad_kernel<W_10>, cudaFuncAttributeMaxDynamicSharedMemorySize, W_10::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(convT2x2_wgrad_kernel<TB_u0>, cudaFuncAttributeMaxDynamicSharedMemorySize, convT_wgrad_smem<TB_u0>()));
  CU_TRY(cudaFuncSetAttribute(convT2x2_wgrad_kernel<TB_u1>, cudaFuncAttributeMaxDynamicSharedMemorySize, convT_wgrad_smem<TB_u1>()));
  CU_TRY(cudaFuncSetAttribute(convT2x2_dgrad_kernel<TB_u0>, cudaFuncAttributeMaxDynamicSharedMemorySize, convT_dgrad_smem<TB_u0>()));
  CU_TRY(cudaFuncSetAttribute(convT2x2_dgrad_kernel<TB_u1>, cudaFuncAttributeMaxDynamicSharedMemorySize, convT_dgrad_smem<TB_u1>()));
  CU_TRY(cudaFuncSetAttribute(decode_points_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(conv_in_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(conv_in_planes_train_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvInCfg<5, 1>::SMEM_BYTES));
  CU_TRY(cudaFuncSetAttribute(conv_in_planes_train_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvInCfg<1, 4>::SMEM_BYTES));
  done[ctx->device & 63] = true;
  return GIGA_OK;
}

// (re)build the packing table for the bound tensors and upload it
int train_build_table(giga_ctx* ctx) {
  auto& T = ctx->tr;
  const EncLayout& L = ctx->el;
  if (!T.d_tab) {
    // the training operands live in the SAME blobs as the inference path's (ctx->d_enc: fp32 prefix shared, data-gradient weights behind the
    // inference layout; ctx->d_heads): one packing pass serves both
    long o = L.total;
    for (int i = 0; i < 10; ++i) { T.dg[i] = o; o += (long)kConvCin[i] * 9 * kConvCout[i]; }
    T.blob_floats = o;
    if (!ctx->d_enc) {
      CU_TRY(cudaMalloc(&ctx->d_enc, sizeof(float) * o));
      CU_TRY(cudaMemset(ctx->d_enc, 0, sizeof(float) * o));
    }
    if (!ctx->d_heads) {
      CU_TRY(cudaMalloc(&ctx->d_heads, sizeof(float) * 4 * DW_HEAD));
      CU_TRY(cudaMemset(ctx->d_heads, 0, sizeof(float) * 4 * DW_HEAD));
    }
    CU_TRY(cudaMalloc(&T.d_cin, sizeof(ConvInParams)));
    CU_TRY(cudaMalloc(&T.d_tab, sizeof(PackEntry) * 256));
  }
  T.d_blob = ctx->d_enc;
  T.d_heads = ctx->d_heads;
  std::vector<PackEntry> tab;
  auto add = [&](const float* src, float* dst, int O, int I, int Tt, int mode, int ld, int i0 = 0, int In = 0) {
    PackEntry e = {src, dst, O, I, Tt, mode, ld, i0, In, 0};
    tab.push_back(e);
  };
  float* E = T.d_blob;
  add(T.val[0], T.d_cin, 32, 1, 27, 0, 32);
  add(T.val[1], T.d_cin + 27 * 32, 32, 1, 1, 2, 0);
  for (int i = 0; i < 10; ++i) {
    const int ci = kConvCin[i], co = kConvCout[i];
    add(T.val[2 + 2 * i], E + L.conv[i], co, ci, 9, 0, co);
    add(T.val[3 + 2 * i], E + L.bias[i], co, 1, 1, 2, 0);
    if (i == 6 || i == 8) {   // concat layers: one data-gradient block per source (upsampled | skip)
      const int half = ci / 2;
      add(T.val[2 + 2 * i], E + T.dg[i], co, ci, 9, 1, 0, 0, half);
      add(T.val[2 + 2 * i], E + T.dg[i] + (long)co * 9 * half, co, ci, 9, 1, 0, half, half);
    } else {
      add(T.val[2 + 2 * i], E + T.dg[i], co, ci, 9, 1, 0, 0, ci);
    }
  }
  for (int i = 0; i < 2; ++i) {
    add(T.val[22 + 2 * i], E + L.up_w[i], kUpCin[i] * kUpCout[i] * 4, 1, 1, 2, 0);
    add(T.val[23 + 2 * i], E + L.up_b[i], kUpCout[i], 1, 1, 2, 0);
  }
  add(T.val[26], E + L.fin_w, 32, 32, 1, 0, 32);
  add(T.val[27], E + L.fin_b, 32, 1, 1, 2, 0);
  for (int h = 0; h < 4; ++h) {
    if (!(T.heads & (1u << h))) continue;
    const float* const* v = T.val + TS_HEAD0 + TS_HEAD * h;
    float* H = T.d_heads + (size_t)h * DW_HEAD;
    const int od = h == 1 ? 4 : 1;
    add(v[0], H + DW_FCP, 32, 3, 1, 0, 32);
    add(v[1], H + DW_FCP + 96, 32, 1, 1, 2, 0);
    for (int i = 0; i < 5; ++i) {
      float* Bk = H + DW_BLOCK0 + i * DW_BLK;
      add(v[2 + 6 * i], Bk + DW_BLK_FCC, 32, 96, 1, 0, 32);
      add(v[3 + 6 * i], Bk + DW_BLK_BC, 32, 1, 1, 2, 0);
      add(v[4 + 6 * i], Bk + DW_BLK_W0, 32, 32, 1, 0, 32);
      add(v[5 + 6 * i], Bk + DW_BLK_B0, 32, 1, 1, 2, 0);
      add(v[6 + 6 * i], Bk + DW_BLK_W1, 32, 32, 1, 0, 32);
      add(v[7 + 6 * i], Bk + DW_BLK_B1, 32, 1, 1, 2, 0);
    }
    add(v[32], H + DW_OUT, od, 32, 1, 0, 4);
    add(v[33], H + DW_OUT + 128, od, 1, 1, 2, 0);
  }
  if (tab.size() > 256) return fail(GIGA_ESTATE, "giga_train: packing table overflow");
  CU_TRY(cudaDeviceSynchronize());   // a previous step's packing kernel may still read the table
  CU_TRY(cudaMemcpy(T.d_tab, tab.data(), sizeof(PackEntry) * tab.size(), cudaMemcpyHostToDevice));
  T.n_tab = (int)tab.size();
  T.table_dirty = false;
  return GIGA_OK;
}

int train_workspace(giga_ctx* ctx, int B) {
  auto& T = ctx->tr;
  if (B <= T.cap_B) return GIGA_OK;
  CU_TRY(cudaDeviceSynchronize());
  for (auto& p : T.d_g) { if (p) cudaFree(p); p = nullptr; }
  float** one[] = {&T.d_gpre, &T.d_gplanes, &T.d_planes};
  for (float** q : one) { if (*q) cudaFree(*q); *q = nullptr; }
  T.cap_B = 0;
  for (int i = 0; i < kNumActs; ++i) CU_TRY(cudaMalloc(&T.d_g[i], sizeof(float) * 3 * (size_t)B * kActs[i].ch * kActs[i].hw * kActs[i].hw));
  CU_TRY(cudaMalloc(&T.d_gpre, sizeof(float) * 3 * (size_t)B * C * G2));
  CU_TRY(cudaMalloc(&T.d_gplanes, sizeof(float) * 3 * (size_t)B * C * G2));
  CU_TRY(cudaMalloc(&T.d_planes, sizeof(float) * 3 * (size_t)B * C * G2));
  T.cap_B = B;
  return GIGA_OK;
}

float* gact(giga_ctx* ctx, const char* name) {
  for (int i = 0; i < kNumActs; ++i)
    if (!strcmp(kActs[i].name, name)) return ctx->tr.d_g[i];
  return nullptr;
}

template <class K>
void launch_dgrad(giga_ctx* ctx, const char* name, int n_img, const float* gz, const float* w, const float* mask, float* out, cudaStream_t st) {
  dim3 grid(K::NB * K::NCT, n_img);
  LaunchScope ls(ctx, name, st);
  conv3x3_kernel<K><<<grid, K::NTHREADS, K::SMEM_BYTES, st>>>(gz, nullptr, w, mask, out, nullptr);
}

template <class K>
void launch_wgrad(giga_ctx* ctx, const char* name, int n_img, const float* gz, const float* in, int cout, int cin_src, float* dW, int cin_total,
                  float* db, cudaStream_t st) {
  const int n_cc = (cout / 32) * (cin_src / 32), n_tiles = n_img * K::NB;
  const int P = std::max(1, std::min(n_tiles, (3 * ctx->num_sms + n_cc - 1) / n_cc));
  unsigned* counters = ctx->tr.d_wg_counters + 16 * (ctx->tr.wg_slot++ % 16);   // zeroed at the start of the backward
  LaunchScope ls(ctx, name, st);
  conv3x3_wgrad_kernel<K><<<dim3(n_cc, P), 256, K::SMEM_BYTES, st>>>(gz, in, n_img, cout, cin_src, dW, cin_total, db, counters);
}

template <class K>
void launch_convT_bwd(giga_ctx* ctx, const char* name, int n_img, const float* in, const float* gout, const float* w, float* gin, float* dW, float* db,
                      cudaStream_t st) {
  {
    const int n_cc = (K::CIN / 32) * (K::COUT / 32), n_tiles = n_img * K::NB;
    const int P = std::max(1, std::min(n_tiles, (2 * ctx->num_sms + n_cc - 1) / n_cc));
    LaunchScope ls(ctx, name, st);
    convT2x2_wgrad_kernel<K><<<dim3(n_cc, P), 256, convT_wgrad_smem<K>(), st>>>(in, gout, n_img, dW, db);
  }
  {
    LaunchScope ls(ctx, name, st);
    convT2x2_dgrad_kernel<K><<<dim3(K::NB * (K::CIN / 32), n_img), 256, convT_dgrad_smem<K>(), st>>>(gout, w, in, gin);
  }
}

// names -> slots; validates completeness (all encoder tensors, whole heads)
int map_param_slots(const char* what, int n, const char* const* names, const float* const* values, float* const* grads,
                    const float** val, float** grad, unsigned* heads_out) {
  static std::map<std::string, int> slot;
  if (slot.empty())
    for (int s = 0; s < giga_ctx::Train::kSlots; ++s) slot[train_slot_name(s)] = s;
  for (int i = 0; i < n; ++i) {
    if (!names[i] || !values[i]) return fail(GIGA_EINVAL, std::string(what) + ": null name or tensor");
    auto it = slot.find(names[i]);
    if (it == slot.end()) return fail(GIGA_EINVAL, std::string(what) + ": unknown parameter '" + names[i] + "'");
    if ((reinterpret_cast<uintptr_t>(values[i]) & 15) || (grads && (reinterpret_cast<uintptr_t>(grads[i]) & 15)))
      return fail(GIGA_EINVAL, std::string(what) + ": '" + names[i] + "' is not 16-byte aligned");
    val[it->second] = values[i];
    if (grads) grad[it->second] = grads[i];
  }
  for (int s = 0; s < 28; ++s)
    if (!val[s]) return fail(GIGA_ESTATE, std::string(what) + ": missing encoder parameter '" + train_slot_name(s) + "'");
  unsigned heads = 0;
  for (int h = 0; h < 4; ++h) {
    int have = 0;
    for (int r = 0; r < TS_HEAD; ++r) have += val[TS_HEAD0 + TS_HEAD * h + r] != nullptr;
    if (have == TS_HEAD) heads |= 1u << h;
    else if (have) return fail(GIGA_ESTATE, std::string(what) + ": decoder_" + kHeadName[h] + " is incomplete");
  }
  if (!heads) return fail(GIGA_ESTATE, std::string(what) + ": no decoder head");
  *heads_out = heads;
  return GIGA_OK;
}

template <class K>
TcPackEntry tc_entry(const float* src, float* dst_words) {
  TcPackEntry e = {src, reinterpret_cast<uint16_t*>(dst_words), K::NNT, K::NC, K::NTAPS, K::NTILE, K::CIN, K::COUT, K::MODE, 0};
  return e;
}

// blobs + tables of the device-side packer (pack_dev.cuh) for the bound tensors
int ensure_tc_tables(giga_ctx* ctx) {
  auto& T = ctx->tr;
  const EncLayout& L = ctx->el;
  float* E = ctx->d_enc;
  if (!ctx->d_hc) {
    CU_TRY(cudaMalloc(&ctx->d_hc, sizeof(float) * 4 * HC_SIZE));
    CU_TRY(cudaMemset(ctx->d_hc, 0, sizeof(float) * 4 * HC_SIZE));
  }
  if (!ctx->d_sched) {
    CU_TRY(cudaMalloc(&ctx->d_sched, sizeof(unsigned) * 4));
    CU_TRY(cudaMemset(ctx->d_sched, 0, sizeof(unsigned) * 4));
  }
  for (int type = 0; type < 5; ++type) {
    const int nh = type == 0 ? 3 : 1;
    const unsigned need = type == 0 ? 7u : (1u << (type - 1));
    if ((T.heads & need) != need) {
      if (ctx->d_wblob[type]) { CU_TRY(cudaDeviceSynchronize()); cudaFree(ctx->d_wblob[type]); ctx->d_wblob[type] = nullptr; }
    } else if (!ctx->d_wblob[type]) {
      CU_TRY(cudaMalloc(&ctx->d_wblob[type], (size_t)5 * wd_block_bytes(nh)));
      CU_TRY(cudaMemset(ctx->d_wblob[type], 0, (size_t)5 * wd_block_bytes(nh)));
    }
  }
  if (!T.d_tctab) {
    CU_TRY(cudaMalloc(&T.d_tctab, sizeof(TcPackEntry) * 16));
    CU_TRY(cudaMalloc(&T.d_hscale, sizeof(float) * 8));
  }
  if (T.tc_table_dirty) {
    std::vector<TcPackEntry> tab;
    auto cw = [&](int i) { return T.val[2 + 2 * i]; };
    tab.push_back(tc_entry<T_c40>(cw(0), E + L.tc_conv[0]));
    tab.push_back(tc_entry<T_c40>(cw(1), E + L.tc_conv[1]));
    tab.push_back(tc_entry<T_d1c1>(cw(2), E + L.tc_conv[2]));
    tab.push_back(tc_entry<T_c20>(cw(3), E + L.tc_conv[3]));
    tab.push_back(tc_entry<T_d2c1>(cw(4), E + L.tc_conv[4]));
    tab.push_back(tc_entry<T_d2c2>(cw(5), E + L.tc_conv[5]));
    tab.push_back(tc_entry<T_u0c1>(cw(6), E + L.tc_conv[6]));
    tab.push_back(tc_entry<T_c20>(cw(7), E + L.tc_conv[7]));
    tab.push_back(tc_entry<T_u1c1>(cw(8), E + L.tc_conv[8]));
    tab.push_back(tc_entry<T_u1c2>(cw(9), E + L.tc_conv[9]));
    tab.push_back(tc_entry<T_u0up>(T.val[22], E + L.tc_up[0]));
    tab.push_back(tc_entry<T_u1up>(T.val[24], E + L.tc_up[1]));
    TcPackEntry fin = {T.val[26], reinterpret_cast<uint16_t*>(E + L.tc_fin), 0, 0, 0, 0, 0, 0, 2, 0};
    tab.push_back(fin);
    CU_TRY(cudaDeviceSynchronize());   // a packing kernel of an earlier commit may still read the table
    CU_TRY(cudaMemcpy(T.d_tctab, tab.data(), sizeof(TcPackEntry) * tab.size(), cudaMemcpyHostToDevice));
    T.n_tctab = (int)tab.size();
    T.tc_table_dirty = false;
  }
  return GIGA_OK;
}

DecPackArgs make_dec_pack_args(giga_ctx* ctx) {
  auto& T = ctx->tr;
  DecPackArgs A = {};
  A.heads = T.heads;
  for (int h = 0; h < 4; ++h) {
    if (!(T.heads & (1u << h))) continue;
    const float* const* v = T.val + TS_HEAD0 + TS_HEAD * h;
    for (int i = 0; i < 5; ++i) {
      A.head[h].fcc_w[i] = v[2 + 6 * i]; A.head[h].fcc_b[i] = v[3 + 6 * i];
      A.head[h].w0[i] = v[4 + 6 * i];   A.head[h].b0[i] = v[5 + 6 * i];
      A.head[h].w1[i] = v[6 + 6 * i];   A.head[h].b1[i] = v[7 + 6 * i];
    }
  }
  for (int t = 0; t < 5; ++t) A.wblob[t] = ctx->d_wblob[t];
  A.hc = ctx->d_hc;
  A.dheads = ctx->d_heads;
  A.scale = T.d_hscale;
  return A;
}

// all operand layouts from the live tensors: fp32 blobs (+ data-gradient weights, conv_in staging), tensor-core convs, decoder
void launch_device_pack(giga_ctx* ctx, const DecPackArgs& A, cudaStream_t st) {
  auto& T = ctx->tr;
  {
    LaunchScope ls(ctx, "pack:fp32", st);
    train_pack_kernel<<<dim3(T.n_tab, 4), 256, 0, st>>>(T.d_tab);
  }
  {
    LaunchScope ls(ctx, "pack:tc_conv", st);
    pack_conv_tc_kernel<<<dim3(T.n_tctab, 8), 256, 0, st>>>(T.d_tctab);
  }
  {
    LaunchScope ls(ctx, "pack:head_scale", st);
    head_scale_kernel<<<4, 256, 0, st>>>(A);
  }
  {
    LaunchScope ls(ctx, "pack:decoder", st);
    pack_decoder_ws_kernel<<<dim3(5, 5), 256, 0, st>>>(A);
  }
}

}  // namespace

extern "C" {

int giga_ctx_commit_device(giga_ctx* ctx, int n, const char* const* names, const float* const* values, void* stream) {
  if (!ctx || n <= 0 || !names || !values) return fail(GIGA_EINVAL, "giga_ctx_commit_device: bad argument");
  if (int r = set_device(ctx)) return r;
  auto& T = ctx->tr;
  const float* val[giga_ctx::Train::kSlots] = {};
  unsigned heads = 0;
  if (int r = map_param_slots("giga_ctx_commit_device", n, names, values, nullptr, val, nullptr, &heads)) return r;
  bool same = T.bound && heads == T.heads;
  for (int s = 0; same && s < giga_ctx::Train::kSlots; ++s) same = val[s] == T.val[s];
  if (!same) {
    memcpy(T.val, val, sizeof val);
    memset(T.grad, 0, sizeof T.grad);   // gradient buffers belong to a giga_train_bind of the same tensors
    T.heads = heads;
    T.bound = true;
    T.table_dirty = T.tc_table_dirty = true;
  }
  if (int r = ensure_attrs(ctx)) return r;
  // the packed blobs may still be read by kernels of this ctx on other streams (torch side streams, the pipelined host path)
  CU_TRY(cudaDeviceSynchronize());
  if (T.table_dirty)
    if (int r = train_build_table(ctx)) return r;
  if (int r = ensure_tc_tables(ctx)) return r;
  DecPackArgs A = make_dec_pack_args(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  launch_device_pack(ctx, A, st);
  // conv_in's 27 x 32 taps are kernel parameters of the inference kernel (constant bank): 3.6 KB come back to the host
  {
    float cin[28 * 32];
    CU_TRY(cudaMemcpyAsync(cin, T.d_cin, sizeof cin, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    memcpy(&ctx->conv_in.w[0][0].x, cin, sizeof(float) * 27 * 32);
    memcpy(&ctx->conv_in.b[0].x, cin + 27 * 32, sizeof(float) * 32);
  }
  CU_TRY(cudaGetLastError());
  ctx->has_encoder = true;
  ctx->has_vgn = false;
  ctx->heads = heads;
  ctx->committed = true;
  ctx->graph_epoch++;
  return GIGA_OK;
}

long giga_debug_blob(giga_ctx* ctx, int which, void* dst, long capacity) {
  if (!ctx || !dst || which < 0 || which > 7) return fail(GIGA_EINVAL, "giga_debug_blob: bad argument");
  if (int r = set_device(ctx)) return r;
  const void* src = nullptr;
  long bytes = 0;
  if (which == 0) { src = ctx->d_enc; bytes = sizeof(float) * ctx->el.total; }
  else if (which == 1) { src = ctx->d_heads; bytes = sizeof(float) * 4 * DW_HEAD; }
  else if (which == 2) { src = ctx->d_hc; bytes = sizeof(float) * 4 * HC_SIZE; }
  else { src = ctx->d_wblob[which - 3]; bytes = (long)5 * wd_block_bytes(which == 3 ? 3 : 1); }
  if (!src) return 0;
  if (bytes > capacity) return fail(GIGA_EINVAL, "giga_debug_blob: capacity too small");
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return bytes;
}

int giga_train_bind(giga_ctx* ctx, int n, const char* const* names, const float* const* values, float* const* grads) {
  if (!ctx || n <= 0 || !names || !values || !grads) return fail(GIGA_EINVAL, "giga_train_bind: bad argument");
  if (int r = set_device(ctx)) return r;
  auto& T = ctx->tr;
  const float* val[giga_ctx::Train::kSlots] = {};
  float* grad[giga_ctx::Train::kSlots] = {};
  unsigned heads = 0;
  if (int r = map_param_slots("giga_train_bind", n, names, values, grads, val, grad, &heads)) return r;
  bool same = T.bound && heads == T.heads;
  for (int s = 0; same && s < giga_ctx::Train::kSlots; ++s) same = val[s] == T.val[s];
  memcpy(T.val, val, sizeof val);
  memcpy(T.grad, grad, sizeof grad);
  T.heads = heads;
  T.bound = true;
  if (!same) T.table_dirty = T.tc_table_dirty = true;
  return GIGA_OK;
}

int giga_train_forward(giga_ctx* ctx, const float* tsdf, int B, const float* p, int Ng, const float* p_tsdf, int No, int detach_tsdf, float* qual,
                       float* rot, float* width, float* occ, void* stream) {
  if (!ctx || !tsdf || B <= 0) return fail(GIGA_EINVAL, "giga_train_forward: bad argument");
  auto& T = ctx->tr;
  if (!T.bound) return fail(GIGA_ESTATE, "giga_train_forward: no parameters bound (giga_train_bind)");
  const bool grasp = p && Ng > 0, geo = p_tsdf && No > 0;
  if (!grasp && !geo) return fail(GIGA_EINVAL, "giga_train_forward: no query points");
  const unsigned hm = T.heads & 7u;
  if (grasp && (hm != 7u || !qual || !rot || !width)) return fail(GIGA_EINVAL, "giga_train_forward: grasp points need the three grasp heads and their outputs");
  if (geo && (!(T.heads & 8u) || !occ)) return fail(GIGA_EINVAL, "giga_train_forward: p_tsdf needs the TSDF head and its output");
  if (int r = set_device(ctx)) return r;
  if (int r = ensure_attrs(ctx)) return r;
  if (int r = train_attrs(ctx)) return r;
  if (T.table_dirty)
    if (int r = train_build_table(ctx)) return r;
  if (int r = ensure_workspace(ctx, B)) return r;
  if (int r = train_workspace(ctx, B)) return r;
  if (int r = ensure_tsdf_maps(ctx, tsdf, B)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  T.fwd_valid = false;
  // ---- operands from the live parameters (device to device; nothing crosses the host) ----
  const bool tc = T.fwd_impl == 1;
  if (tc) {
    if (int r = ensure_tc_tables(ctx)) return r;
    launch_device_pack(ctx, make_dec_pack_args(ctx), st);
  } else {
    LaunchScope ls(ctx, "train:pack", st);
    train_pack_kernel<<<dim3(T.n_tab, 4), 256, 0, st>>>(T.d_tab);
  }
  CU_TRY(cudaMemcpyToSymbolAsync(c_conv_in_train, T.d_cin, sizeof(ConvInParams), 0, cudaMemcpyDeviceToDevice, st));
  // ---- encoder: the inference conv_in body (weights from constant memory), then the U-Net on the tcgen05 kernels (default) or on the fp32
  //      FMA-pipe kernels; every activation is kept (NCHW fp32) for the backward ----
  const int n_img = 3 * B;
  float* tall_pre = ctx->d_tall[0];
  const long ps_pre = ctx->tall_ps[0];
  {
    LaunchScope ls(ctx, "train:conv_in", st);
    if (conv_in_ty(B) == 1) {
      using Cf = ConvInCfg<1, 4>;
      conv_in_planes_train_kernel<1, 4><<<dim3(Cf::NT, B, 4), Cf::THREADS, Cf::SMEM_BYTES, st>>>(ctx->tsdf_map[1], tall_pre, ps_pre, ctx->d_xzpart, B);
    } else {
      using Cf = ConvInCfg<5, 1>;
      conv_in_planes_train_kernel<5, 1><<<dim3(Cf::NT, B, 1), Cf::THREADS, Cf::SMEM_BYTES, st>>>(ctx->tsdf_map[0], tall_pre, ps_pre, ctx->d_xzpart, B);
    }
  }
  {
    LaunchScope ls(ctx, "train:xz_finish", st);
    const int blocks = ceil_div((int)std::max((long)B * 4 * G2, 9 * ctx->flags_stride + 16), 256);
    if (conv_in_ty(B) == 5)
      xz_finish_tall_kernel<G / 5><<<blocks, 256, 0, st>>>((const float*)ctx->d_xzpart, tall_pre, ps_pre, B, ctx->d_flags, (int)(9 * ctx->flags_stride + 16));
    else
      xz_finish_tall_kernel<G><<<blocks, 256, 0, st>>>((const float*)ctx->d_xzpart, tall_pre, ps_pre, B, ctx->d_flags, (int)(9 * ctx->flags_stride + 16));
  }
  ctx->work_slot = 0;
  {
    LaunchScope ls(ctx, "train:tall_to_nchw", st);
    tall_to_nchw_kernel<40, 4><<<ceil_div(n_img * 4 * G2, 256), 256, 0, st>>>(tall_pre, ctx->d_pre, ps_pre, n_img);
  }
  if (tc) {
    // ---- tcgen05 U-Net (the inference kernels; the last layer unfused so that u1c2 exists), then every activation expanded to the NCHW fp32
    //      form the backward kernels read (hi + lo * 2^-11 carries 22 significant bits), conv_final on the expanded u1c2 ----
    unet_tall_forward(ctx, n_img, nullptr, /*keep_u1c2=*/true, st);
    {
      ExpandArgs X = {};
      X.n_img = n_img;
      int max_blocks = 0;
      for (int i = 0; i < kNumActs; ++i) {
        X.e[i] = {ctx->d_tall[1 + i], ctx->d_act[i], ctx->tall_ps[1 + i], kActs[i].hw, kActs[i].ch / 8};
        max_blocks = std::max(max_blocks, ceil_div(n_img * (kActs[i].ch / 8) * kActs[i].hw * kActs[i].hw, 256));
      }
      LaunchScope ls(ctx, "train:expand_activations", st);
      tall_expand_all_kernel<<<dim3(max_blocks, kNumActs), 256, 0, st>>>(X);
    }
    {
      LaunchScope ls(ctx, "train:conv_final", st);
      conv1x1_nhwc_kernel<<<dim3(G2 / F_PIX, n_img), 256, 0, st>>>(act(ctx, "u1c2"), ctx->d_enc + ctx->el.fin_w, ctx->d_enc + ctx->el.fin_b, T.d_planes);
    }
    ctx->last_B = B;
    ctx->last_impl = 0;   // d_act holds every activation in NCHW fp32 (giga_debug_copy reads them there)
    // ---- heads: the warp-specialised tcgen05 decoder on the device-packed blobs, both point sets in one launch ----
    if (grasp && geo) {
      if (int r = launch_decode(ctx, T.d_planes, B, p, Ng, 7u, p_tsdf, No, 8u, qual, rot, width, occ, st)) return r;
    } else if (grasp) {
      if (int r = launch_decode(ctx, T.d_planes, B, p, Ng, 7u, nullptr, 0, 0u, qual, rot, width, nullptr, st)) return r;
    } else {
      if (int r = launch_decode(ctx, T.d_planes, B, p_tsdf, No, 8u, nullptr, 0, 0u, nullptr, nullptr, nullptr, occ, st)) return r;
    }
    CU_TRY(cudaGetLastError());
    T.x = tsdf; T.p = grasp ? p : nullptr; T.pt = geo ? p_tsdf : nullptr;
    T.B = B; T.Ng = grasp ? Ng : 0; T.No = geo ? No : 0; T.detach = detach_tsdf ? 1 : 0;
    T.fwd_valid = true;
    return GIGA_OK;
  }
  float *d0c1 = act(ctx, "d0c1"), *d0c2 = act(ctx, "d0c2"), *p0 = act(ctx, "p0"), *d1c1 = act(ctx, "d1c1"), *d1c2 = act(ctx, "d1c2"),
        *p1 = act(ctx, "p1"), *d2c1 = act(ctx, "d2c1"), *d2c2 = act(ctx, "d2c2"), *u0 = act(ctx, "u0"), *u0c1 = act(ctx, "u0c1"),
        *u0c2 = act(ctx, "u0c2"), *u1 = act(ctx, "u1"), *u1c1 = act(ctx, "u1c1"), *u1c2 = act(ctx, "u1c2");
  const float* E = T.d_blob;
  const EncLayout& L = ctx->el;
  auto conv = [&](auto kcfg, const char* name, const float* s0, const float* s1, int layer, float* out, float* pooled) {
    using K = decltype(kcfg);
    dim3 grid(K::NB * K::NCT, n_img);
    LaunchScope ls(ctx, name, st);
    conv3x3_kernel<K><<<grid, K::NTHREADS, K::SMEM_BYTES, st>>>(s0, s1, E + L.conv[layer], E + L.bias[layer], out, pooled);
  };
  conv(K_d0c1{}, "train:conv:d0c1", ctx->d_pre, nullptr, 0, d0c1, nullptr);
  conv(K_d0c2{}, "train:conv:d0c2", d0c1, nullptr, 1, d0c2, p0);
  conv(K_d1c1{}, "train:conv:d1c1", p0, nullptr, 2, d1c1, nullptr);
  conv(K_d1c2{}, "train:conv:d1c2", d1c1, nullptr, 3, d1c2, p1);
  conv(K_d2c1{}, "train:conv:d2c1", p1, nullptr, 4, d2c1, nullptr);
  conv(K_d2c2{}, "train:conv:d2c2", d2c1, nullptr, 5, d2c2, nullptr);
  {
    LaunchScope ls(ctx, "train:convT:u0", st);
    convT2x2_kernel<K_u0up><<<dim3(K_u0up::NB * K_u0up::NCT, n_img), K_u0up::NTHREADS, K_u0up::SMEM_BYTES, st>>>(d2c2, E + L.up_w[0], E + L.up_b[0], u0);
  }
  conv(K_u0c1{}, "train:conv:u0c1", u0, d1c2, 6, u0c1, nullptr);
  conv(K_u0c2{}, "train:conv:u0c2", u0c1, nullptr, 7, u0c2, nullptr);
  {
    LaunchScope ls(ctx, "train:convT:u1", st);
    convT2x2_kernel<K_u1up><<<dim3(K_u1up::NB * K_u1up::NCT, n_img), K_u1up::NTHREADS, K_u1up::SMEM_BYTES, st>>>(u0c2, E + L.up_w[1], E + L.up_b[1], u1);
  }
  conv(K_u1c1{}, "train:conv:u1c1", u1, d0c2, 8, u1c1, nullptr);
  conv(K_u1c2{}, "train:conv:u1c2", u1c1, nullptr, 9, u1c2, nullptr);
  {
    LaunchScope ls(ctx, "train:conv_final", st);
    conv1x1_nhwc_kernel<<<dim3(G2 / F_PIX, n_img), 256, 0, st>>>(u1c2, E + L.fin_w, E + L.fin_b, T.d_planes);
  }
  ctx->last_B = B;
  ctx->last_impl = 0;
  // ---- heads (fp32 FMA-pipe decoder on the device-packed parameters) ----
  if (grasp) {
    LaunchScope ls(ctx, "train:decode:grasp", st);
    decode_points_kernel<<<dim3(ceil_div(Ng, DEC_PTS), B), DEC_PTS, DEC_SMEM_BYTES, st>>>(T.d_planes, p, T.d_heads, B, Ng, 7u, qual, rot, width, nullptr);
  }
  if (geo) {
    LaunchScope ls(ctx, "train:decode:tsdf", st);
    decode_points_kernel<<<dim3(ceil_div(No, DEC_PTS), B), DEC_PTS, DEC_SMEM_BYTES, st>>>(T.d_planes, p_tsdf, T.d_heads, B, No, 8u, nullptr, nullptr, nullptr, occ);
  }
  CU_TRY(cudaGetLastError());
  T.x = tsdf; T.p = grasp ? p : nullptr; T.pt = geo ? p_tsdf : nullptr;
  T.B = B; T.Ng = grasp ? Ng : 0; T.No = geo ? No : 0; T.detach = detach_tsdf ? 1 : 0;
  T.fwd_valid = true;
  return GIGA_OK;
}

int giga_train_backward(giga_ctx* ctx, const float* g_qual, const float* g_rot, const float* g_width, const float* g_occ, float* g_p, void* stream) {
  if (!ctx) return fail(GIGA_EINVAL, "giga_train_backward: ctx is null");
  auto& T = ctx->tr;
  if (!T.fwd_valid) return fail(GIGA_ESTATE, "giga_train_backward: no forward to differentiate (giga_train_forward must precede; one backward per forward)");
  if ((g_qual || g_rot || g_width) && !T.p) return fail(GIGA_EINVAL, "giga_train_backward: grasp-head gradients without grasp points in the forward");
  if (g_occ && !T.pt) return fail(GIGA_EINVAL, "giga_train_backward: TSDF-head gradient without p_tsdf in the forward");
  if (g_p && !T.p) return fail(GIGA_EINVAL, "giga_train_backward: position gradient without grasp points in the forward");
  if (g_rot && (reinterpret_cast<uintptr_t>(g_rot) & 15)) return fail(GIGA_EINVAL, "giga_train_backward: g_rot must be 16-byte aligned");
  for (int s = 0; s < giga_ctx::Train::kSlots; ++s)
    if (T.val[s] && !T.grad[s]) return fail(GIGA_ESTATE, "giga_train_backward: '" + train_slot_name(s) + "' has no gradient buffer bound");
  if (int r = set_device(ctx)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  OrderScope order(ctx, st);
  T.fwd_valid = false;
  const int B = T.B, n_img = 3 * B;
  const float* gouts[4] = {g_qual, g_rot, g_width, g_occ};
  bool enc_grad = false;
  for (int h = 0; h < 4; ++h) enc_grad |= gouts[h] && !(h == 3 && T.detach);
  if (enc_grad) CU_TRY(cudaMemsetAsync(T.d_gplanes, 0, sizeof(float) * 3 * (size_t)B * C * G2, st));
  if (!T.d_wg_counters) CU_TRY(cudaMalloc(&T.d_wg_counters, sizeof(unsigned) * 256));
  CU_TRY(cudaMemsetAsync(T.d_wg_counters, 0, sizeof(unsigned) * 256, st));   // work counters of the 14 filter-gradient launches
  T.wg_slot = 0;
  // ---- heads: ONE launch, blockIdx.z = (head with a gradient), each at its own point set ----
  {
    DecBwdArgs A = {};
    int nj = 0, max_tiles = 0;
    size_t need = 0;
    for (int h = 0; h < 4; ++h) {
      if (!gouts[h]) continue;
      DecBwdJob& J = A.job[nj++];
      J.head = h;
      J.pts = h == 3 ? T.pt : T.p;
      J.N = h == 3 ? T.No : T.Ng;
      J.tiles = ceil_div(J.N, DB_PTS);
      J.gout = gouts[h];
      J.gplanes = (h == 3 && T.detach) ? nullptr : T.d_gplanes;
      J.gpts = h == 3 ? nullptr : g_p;
      J.save_off = (long)need;
      need += (size_t)B * J.tiles * DB_SAVE;
      max_tiles = std::max(max_tiles, J.tiles);
      float* const* g = T.grad + TS_HEAD0 + TS_HEAD * h;
      J.GR.fcp_w = g[0]; J.GR.fcp_b = g[1];
      for (int i = 0; i < 5; ++i) {
        J.GR.fcc_w[i] = g[2 + 6 * i]; J.GR.fcc_b[i] = g[3 + 6 * i];
        J.GR.w0[i] = g[4 + 6 * i]; J.GR.b0[i] = g[5 + 6 * i];
        J.GR.w1[i] = g[6 + 6 * i]; J.GR.b1[i] = g[7 + 6 * i];
      }
      J.GR.out_w = g[32]; J.GR.out_b = g[33];
    }
    if (nj) {
      if (need > T.save_cap) {
        CU_TRY(cudaStreamSynchronize(st));
        if (T.d_save) cudaFree(T.d_save);
        T.d_save = nullptr;
        T.save_cap = 0;
        CU_TRY(cudaMalloc(&T.d_save, sizeof(float) * need));
        T.save_cap = need;
      }
      LaunchScope ls(ctx, "train:decode_bwd", st);
      decode_points_bwd_kernel<<<dim3(max_tiles, B, nj), DB_PTS, DB_SMEM_BYTES, st>>>(T.d_planes, T.d_heads, B, A, T.d_save);
    }
  }
  CU_TRY(cudaGetLastError());
  if (!enc_grad) return GIGA_OK;
  // ---- encoder ----
  float *d0c1 = act(ctx, "d0c1"), *d0c2 = act(ctx, "d0c2"), *p0 = act(ctx, "p0"), *d1c1 = act(ctx, "d1c1"), *d1c2 = act(ctx, "d1c2"),
        *p1 = act(ctx, "p1"), *d2c1 = act(ctx, "d2c1"), *d2c2 = act(ctx, "d2c2"), *u0 = act(ctx, "u0"), *u0c1 = act(ctx, "u0c1"),
        *u0c2 = act(ctx, "u0c2"), *u1 = act(ctx, "u1"), *u1c1 = act(ctx, "u1c1"), *u1c2 = act(ctx, "u1c2");
  float *g_d0c1 = gact(ctx, "d0c1"), *g_d0c2 = gact(ctx, "d0c2"), *g_p0 = gact(ctx, "p0"), *g_d1c1 = gact(ctx, "d1c1"), *g_d1c2 = gact(ctx, "d1c2"),
        *g_p1 = gact(ctx, "p1"), *g_d2c1 = gact(ctx, "d2c1"), *g_d2c2 = gact(ctx, "d2c2"), *g_u0 = gact(ctx, "u0"), *g_u0c1 = gact(ctx, "u0c1"),
        *g_u0c2 = gact(ctx, "u0c2"), *g_u1 = gact(ctx, "u1"), *g_u1c1 = gact(ctx, "u1c1"), *g_u1c2 = gact(ctx, "u1c2");
  const float* E = T.d_blob;
  const EncLayout& L = ctx->el;
  auto gw = [&](int layer) { return T.grad[2 + 2 * layer]; };
  auto gb = [&](int layer) { return T.grad[3 + 2 * layer]; };
  {
    LaunchScope ls(ctx, "train:conv_final_bwd", st);
    const int n_tiles = n_img * (G2 / FB_PIX);
    conv_final_bwd_kernel<<<std::min(n_tiles, 4 * ctx->num_sms), 256, 0, st>>>(T.d_gplanes, u1c2, E + L.fin_w, g_u1c2, T.grad[26], T.grad[27], n_tiles);
  }
  // u1c2 (layer 9): in = u1c1
  launch_wgrad<W_40>(ctx, "train:wgrad:u1c2", n_img, g_u1c2, u1c1, 32, 32, gw(9), 32, gb(9), st);
  launch_dgrad<G_40>(ctx, "train:dgrad:u1c2", n_img, g_u1c2, E + T.dg[9], u1c1, g_u1c1, st);
  // u1c1 (layer 8): in = cat(u1, d0c2)
  launch_wgrad<W_40>(ctx, "train:wgrad:u1c1", n_img, g_u1c1, u1, 32, 32, gw(8), 64, gb(8), st);
  launch_wgrad<W_40>(ctx, "train:wgrad:u1c1", n_img, g_u1c1, d0c2, 32, 32, gw(8) + 32 * 9, 64, nullptr, st);
  launch_dgrad<G_40>(ctx, "train:dgrad:u1c1", n_img, g_u1c1, E + T.dg[8], nullptr, g_u1, st);
  launch_dgrad<G_40>(ctx, "train:dgrad:u1c1", n_img, g_u1c1, E + T.dg[8] + 32L * 9 * 32, d0c2, g_d0c2, st);
  // upconv 1: u1 = convT(u0c2)
  launch_convT_bwd<TB_u1>(ctx, "train:convT_bwd:u1", n_img, u0c2, g_u1, E + L.up_w[1], g_u0c2, T.grad[24], T.grad[25], st);
  // u0c2 (layer 7): in = u0c1
  launch_wgrad<W_20>(ctx, "train:wgrad:u0c2", n_img, g_u0c2, u0c1, 64, 64, gw(7), 64, gb(7), st);
  launch_dgrad<G_20b>(ctx, "train:dgrad:u0c2", n_img, g_u0c2, E + T.dg[7], u0c1, g_u0c1, st);
  // u0c1 (layer 6): in = cat(u0, d1c2)
  launch_wgrad<W_20>(ctx, "train:wgrad:u0c1", n_img, g_u0c1, u0, 64, 64, gw(6), 128, gb(6), st);
  launch_wgrad<W_20>(ctx, "train:wgrad:u0c1", n_img, g_u0c1, d1c2, 64, 64, gw(6) + 64 * 9, 128, nullptr, st);
  launch_dgrad<G_20b>(ctx, "train:dgrad:u0c1", n_img, g_u0c1, E + T.dg[6], nullptr, g_u0, st);
  launch_dgrad<G_20b>(ctx, "train:dgrad:u0c1", n_img, g_u0c1, E + T.dg[6] + 64L * 9 * 64, d1c2, g_d1c2, st);
  // upconv 0: u0 = convT(d2c2)
  launch_convT_bwd<TB_u0>(ctx, "train:convT_bwd:u0", n_img, d2c2, g_u0, E + L.up_w[0], g_d2c2, T.grad[22], T.grad[23], st);
  // d2c2 (layer 5): in = d2c1;  d2c1 (layer 4): in = p1
  launch_wgrad<W_10>(ctx, "train:wgrad:d2c2", n_img, g_d2c2, d2c1, 128, 128, gw(5), 128, gb(5), st);
  launch_dgrad<G_10b>(ctx, "train:dgrad:d2c2", n_img, g_d2c2, E + T.dg[5], d2c1, g_d2c1, st);
  launch_wgrad<W_10>(ctx, "train:wgrad:d2c1", n_img, g_d2c1, p1, 128, 64, gw(4), 64, gb(4), st);
  launch_dgrad<G_10a>(ctx, "train:dgrad:d2c1", n_img, g_d2c1, E + T.dg[4], nullptr, g_p1, st);
  {
    LaunchScope ls(ctx, "train:pool_bwd:p1", st);
    const long total = (long)n_img * 64 * 100;
    pool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d1c2, g_p1, g_d1c2, 10, total);
  }
  // d1c2 (layer 3): in = d1c1;  d1c1 (layer 2): in = p0
  launch_wgrad<W_20>(ctx, "train:wgrad:d1c2", n_img, g_d1c2, d1c1, 64, 64, gw(3), 64, gb(3), st);
  launch_dgrad<G_20b>(ctx, "train:dgrad:d1c2", n_img, g_d1c2, E + T.dg[3], d1c1, g_d1c1, st);
  launch_wgrad<W_20>(ctx, "train:wgrad:d1c1", n_img, g_d1c1, p0, 64, 32, gw(2), 32, gb(2), st);
  launch_dgrad<G_20a>(ctx, "train:dgrad:d1c1", n_img, g_d1c1, E + T.dg[2], nullptr, g_p0, st);
  {
    LaunchScope ls(ctx, "train:pool_bwd:p0", st);
    const long total = (long)n_img * 32 * 400;
    pool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d0c2, g_p0, g_d0c2, 20, total);
  }
  // d0c2 (layer 1): in = d0c1;  d0c1 (layer 0): in = the pre-U-Net planes
  launch_wgrad<W_40>(ctx, "train:wgrad:d0c2", n_img, g_d0c2, d0c1, 32, 32, gw(1), 32, gb(1), st);
  launch_dgrad<G_40>(ctx, "train:dgrad:d0c2", n_img, g_d0c2, E + T.dg[1], d0c1, g_d0c1, st);
  launch_wgrad<W_40>(ctx, "train:wgrad:d0c1", n_img, g_d0c1, ctx->d_pre, 32, 32, gw(0), 32, gb(0), st);
  launch_dgrad<G_40>(ctx, "train:dgrad:d0c1", n_img, g_d0c1, E + T.dg[0], nullptr, T.d_gpre, st);
  {
    LaunchScope ls(ctx, "train:conv_in_bwd", st);
    conv_in_bwd_kernel<<<std::min(B * G, 2 * ctx->num_sms), CB_THREADS, CB_SMEM_BYTES, st>>>(T.x, T.d_cin, T.d_gpre, B, T.grad[0], T.grad[1]);
  }
  (void)u0; (void)u1; (void)p0; (void)p1;
  CU_TRY(cudaGetLastError());
  return GIGA_OK;
}

}  // extern "C"
