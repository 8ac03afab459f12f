// Yolact handle: model load, execution plan, batched inference, the reference's classify() pipeline
// and the YOLACT detection head.  Replaces Yolact::init / Yolact::classify
// (/root/reference/src/yolact.rs:17-41) and everything below them (classify, classify_tile,
// postprocess, terrible_id, interpreter.invoke()).
//
// Design: the TFLite graph is planned once at create time into a flat list of kernel launches over
// [tile]-batched NHWC buffers.  RESHAPE is an alias, CONCATENATION inputs are produced directly inside
// the concat buffer (tensors carry a byte stride between tiles), PAD is folded into the consuming
// convolution, per-channel requantisation constants / byte LUTs are tabulated on the host with the exact
// TFLite integer rules, and the whole per-batch launch sequence is replayed from a CUDA graph.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <thread>
#include <vector>

#include "common.h"
#include "conv_tc.h"
#include "fixedpoint.cuh"
#include "ops.h"
#include "post.h"
#include "tflite_reader.h"

namespace tod {
namespace {

constexpr int64_t kAlign = 256;
inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

enum StepKind { kStepConvDirect, kStepConvTc, kStepDepthwise, kStepAdd, kStepLut, kStepPad, kStepResize, kStepCopy };

struct Place {   // where a tensor lives at run time
  uint8_t* base = nullptr;   // tile 0
  int64_t tile_stride = 0;   // bytes between tiles
  int64_t bytes = 0;         // stored bytes per tile (pixels * c_store for a channel-padded tensor)
  int c = 0, c_store = 0;    // real / stored channels per pixel (c_store > c: zero-weight padding lanes, see plan())
};

struct Step {
  StepKind kind;
  int op = -1;             // index into Graph::ops
  int in0 = -1, in1 = -1, out = -1;
  ConvGeom g{};
  int32_t in_zp = 0;
  // offsets into the constant arena (-1 = absent)
  int64_t w_off = -1, bias_off = -1, wsum_off = -1, mult_off = -1, shift_off = -1, lut_off = -1, fast_off = -1;
  int32_t out_zp = 0, act_min = 0, act_max = 0;
  AddParams add{};
  int pad_top = 0, pad_left = 0;
  int8_t fill = 0;
  bool align_corners = false, half_pixel = false;
  int64_t copy_dst_off = 0, copy_bytes = 0;
  int64_t macs = 0;
  ConvTc* tc = nullptr;
  std::vector<int> deps;   // indices of the steps whose output this step reads (RESHAPE / CONCAT are transparent)
  std::vector<uint8_t> lut_host;  // kStepLut: the table; conv steps: the composed byte map fused behind the requantisation
  int64_t post_lut_off = -1;
  bool out_moved = false;  // the conv writes where a fused byte-map chain's last output (or a fused ADD's output) lives
  bool fused_add = false;  // `add` holds a residual ADD applied in this conv's epilogue; in1 = the ADD's other input
  bool relu_tab = false;   // the table at fast_off holds the ReLU form (Requant::relu_tab)
  bool add_conv_is_a = true;
  Place out_place;
  // sibling-head fusion: a second, narrow convolution over the same input rides this one's launch (ConvTcArgs::out_oc ...)
  int sib_out = -1;        // tensor the sibling's bytes are written to (its possibly moved output)
  int sib_cols = 0, sib_host_oc = 0;
  int32_t sib_zp = 0;
  Place sib_place;
  int64_t sib_lut_off = -1;
  std::vector<uint8_t> sib_lut_host;
};

void conv_out_pad(int padding, int in, int k, int stride, int dil, int* out, int* pad) {
  const int eff = (k - 1) * dil + 1;
  *out = padding == kSame ? (in + stride - 1) / stride : (in + stride - eff) / stride;
  const int total = std::max(0, (*out - 1) * stride + eff - in);
  *pad = padding == kSame ? total / 2 : 0;
}

void activation_range(int act, const GTensor& out, int32_t* lo, int32_t* hi) {
  const int32_t qmin = out.type == kU8 ? 0 : -128, qmax = out.type == kU8 ? 255 : 127;
  const float scale = out.scale();
  const int32_t zp = out.zp();
  auto quant = [&](float f) { return zp + int32_t(std::round(f / scale)); };
  *lo = qmin;
  *hi = qmax;
  if (act == kActRelu) *lo = std::max(qmin, quant(0.0f));
  else if (act == kActRelu6) {
    *lo = std::max(qmin, quant(0.0f));
    *hi = std::min(qmax, quant(6.0f));
  } else if (act == kActReluN1To1) {
    *lo = std::max(qmin, quant(-1.0f));
    *hi = std::min(qmax, quant(1.0f));
  }
}

struct ConstArena {
  std::vector<uint8_t> host;
  int64_t add(const void* p, size_t bytes) {
    const int64_t off = round_up(int64_t(host.size()), kAlign);
    host.resize(size_t(off) + bytes);
    if (bytes) std::memcpy(host.data() + off, p, bytes);
    return off;
  }
};

// image 0.24.1 imageops::resize sampling weights (Triangle filter, support 1.0); yolact.rs:208,231
struct AxisTables {
  std::vector<int> left, count;
  std::vector<float> weight;
  int in_size = 0, out_size = 0;
  int *d_left = nullptr, *d_count = nullptr;
  float* d_weight = nullptr;
};

int build_axis(int in_size, int out_size, AxisTables* t) {
  t->in_size = in_size;
  t->out_size = out_size;
  t->left.assign(out_size, 0);
  t->count.assign(out_size, 0);
  t->weight.assign(size_t(out_size) * kMaxTaps, 0.f);
  const float ratio = float(in_size) / float(out_size);
  const float sratio = ratio < 1.0f ? 1.0f : ratio;
  const float support = 1.0f * sratio;
  for (int o = 0; o < out_size; ++o) {
    float input = (float(o) + 0.5f) * ratio;
    int64_t left = int64_t(std::floor(input - support));
    left = std::min<int64_t>(std::max<int64_t>(left, 0), in_size - 1);
    int64_t right = int64_t(std::ceil(input + support));
    right = std::min<int64_t>(std::max<int64_t>(right, left + 1), in_size);
    input = input - 0.5f;
    const int n = int(right - left);
    if (n > kMaxTaps) return fail(TOD_ERR_UNSUPPORTED, "resize %d -> %d needs %d taps (max %d)", in_size, out_size, n, kMaxTaps);
    float sum = 0.f;
    float* w = &t->weight[size_t(o) * kMaxTaps];
    for (int i = 0; i < n; ++i) {
      const float a = std::fabs((float(left + i) - input) / sratio);
      w[i] = a < 1.0f ? 1.0f - a : 0.0f;
      sum += w[i];
    }
    for (int i = 0; i < n; ++i) w[i] /= sum;
    t->left[o] = int(left);
    t->count[o] = n;
  }
  return TOD_OK;
}

}  // namespace
}  // namespace tod

using namespace tod;

struct tod_yolact {
  int device = 0;
  int diag_skip = 0;   // TOD_DIAG_SKIP at creation: step classes left out of the pipeline (timing attribution; bench.py's conv-only leg)
  tod_yolact_options opt{};
  Graph graph;
  std::vector<Step> steps;
  std::vector<Place> place;   // per tensor
  std::vector<char> fused_away;  // per tensor: never written because its byte map was fused into the producing conv
  int64_t macs_per_tile = 0;
  int tc_layers = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // early read-back of the tile class maps while detection still runs
  uint32_t* d_tile_bits = nullptr;      // [max_tiles][max_dets][th*tw/32], allocated on the first request for tile-resolution masks
  cudaStream_t nms_stream = nullptr;    // the worst-case Fast-NMS launch runs beside the common-case one
  cudaEvent_t nms_fork = nullptr, nms_join = nullptr;
  cudaEvent_t seg_ready = nullptr;      // recorded (as an external event node) inside the graph right after seg_post_kernel
  bool seg_ready_in_graph = false;
  uint8_t* d_const = nullptr;
  uint8_t* d_act = nullptr;
  size_t act_bytes = 0;
  // literal post-processing
  int seg_out = -1;
  uint32_t* d_tile_classes = nullptr;  // [max_tiles][th][tw]
  uint32_t* d_cell_classes = nullptr;  // [max_tiles][gh][gw]: the same values before the 8x replication of yolact.rs:127-128
  cudaEvent_t caller_done = nullptr;   // recorded on a caller's stream: later fetches on the handle's stream wait for it
  cudaEvent_t done_ev = nullptr, done_ev2 = nullptr;  // blocking-sync events: a waiting host thread sleeps instead of spinning
  bool spin_sync = false;              // TOD_SPIN_SYNC=1: cudaStreamSynchronize (lower wake-up latency, burns a core per waiting thread)
  int* d_diverges = nullptr;
  int* h_diverges = nullptr;           // pinned
  // detection
  int o_box = -1, o_cls = -1, o_coef = -1, o_proto = -1;
  bool det_ready = false, have_priors = false;
  DetectCfg dcfg{};
  DetectBuffers dbuf{};
  std::vector<void*> det_allocs;
  float* d_priors = nullptr;
  // classify() pipeline
  int cw = 0, chh = 0;  // frame size the tables below were built for
  AxisTables pre_v, pre_h, post_v, post_h;
  float* d_tmp = nullptr;
  uint32_t* d_frames = nullptr;
  uint8_t* d_tiles_rgb = nullptr;
  // independent branches of the graph (FPN levels, the five head levels, protonet) are captured on separate
  // streams so the CUDA graph runs them concurrently
  static constexpr int kLanes = 16;
  cudaStream_t lanes[kLanes] = {};
  std::vector<cudaEvent_t> step_events;
  cudaEvent_t post_events[2] = {nullptr, nullptr};  // seg post-processing / detection boxes, when they run on a branch lane
  std::vector<std::vector<int>> out_steps;        // per graph output: the steps that write it
  cudaEvent_t fork_event = nullptr;
  std::vector<int> trace_lanes;
  std::vector<cudaEvent_t> trace_events;  // tod_yolact_trace_steps: timing events, one per step + 4 (start, seg post, boxes, masks)
  // CUDA graphs, keyed by (tiles << 3 | dets << 2 | mask mode); mask mode: 0 = none, 1 = binary masks only, 2 = float + binary
  std::map<int, cudaGraphExec_t> graphs;
  int last_tiles = 0;
  int last_mask_mode = 0;
  int launches_per_call = 0;

  const GTensor& T(int i) const { return graph.tensors[i]; }
  int tile_w() const { return T(graph.inputs[0]).dims[2]; }
  int tile_h() const { return T(graph.inputs[0]).dims[1]; }
};

namespace {

// ------------------------------------------------------------------ planning
int plan(tod_yolact* y, ConstArena* arena) {
  const Graph& G = y->graph;
  const int nt = int(G.tensors.size());
  const bool fuse = y->opt.fusion != 0;

  for (int t : G.outputs)
    if (t < 0 || t >= nt) return fail(TOD_ERR_MODEL, "graph output %d out of range", t);
  if (G.inputs.empty() || G.inputs[0] < 0 || G.inputs[0] >= nt) return fail(TOD_ERR_MODEL, "graph input index out of range");
  for (size_t i = 0; i < G.ops.size(); ++i) {
    const GOp& op = G.ops[i];
    if (op.outputs.empty() || op.inputs.empty()) return fail(TOD_ERR_MODEL, "operator %zu has no inputs or no outputs", i);
    for (int t : op.outputs)
      if (t < 0 || t >= nt) return fail(TOD_ERR_MODEL, "operator %zu: output tensor index %d out of range", i, t);
    for (int t : op.inputs)
      if (t >= nt || t < -1) return fail(TOD_ERR_MODEL, "operator %zu: input tensor index %d out of range", i, t);
    if (op.inputs[0] < 0) return fail(TOD_ERR_MODEL, "operator %zu: first input is absent", i);
    if (op.code == kConcat)
      for (int t : op.inputs)
        if (t < 0) return fail(TOD_ERR_MODEL, "CONCATENATION operator %zu: absent input", i);
    if ((op.code == kConv2D || op.code == kDepthwise || op.code == kPad || op.code == kAdd) && (op.inputs.size() < 2 || op.inputs[1] < 0))
      return fail(TOD_ERR_MODEL, "operator %zu: second input is absent", i);
  }
  for (int t = 0; t < nt; ++t) {   // dimension products must stay far from int64 overflow
    double prod = 1.0;
    for (int d = 0; d < 4; ++d) {
      if (G.tensors[t].dims[d] < 0) return fail(TOD_ERR_MODEL, "tensor %d has a negative dimension", t);
      prod *= double(G.tensors[t].dims[d] > 0 ? G.tensors[t].dims[d] : 1);
    }
    if (prod > 1e12) return fail(TOD_ERR_MODEL, "tensor %d is implausibly large", t);
  }
  const GTensor& in0 = G.tensors[G.inputs[0]];
  if (in0.type != kU8 || in0.dims[0] != 1 || in0.dims[3] != 3)
    return fail(TOD_ERR_MODEL, "expected a uint8 [1,H,W,3] input (yolact.rs:143-153), got type %d [%d,%d,%d,%d]", in0.type, in0.dims[0],
                in0.dims[1], in0.dims[2], in0.dims[3]);

  std::vector<int> producer(nt, -1), consumers(nt, 0);
  for (size_t i = 0; i < G.ops.size(); ++i) {
    for (int t : G.ops[i].outputs) producer[t] = int(i);
    for (int t : G.ops[i].inputs)
      if (t >= 0) consumers[t]++;
  }
  for (int t : G.outputs) consumers[t]++;

  // ---- storage: RESHAPE aliases its input; CONCATENATION inputs are placed inside the output
  std::vector<int> alias_of(nt, -1);
  struct Parent { int tensor = -1; int64_t off = 0; };
  std::vector<Parent> parent(nt);
  auto storage = [&](int t) {
    while (alias_of[t] >= 0) t = alias_of[t];
    return t;
  };
  for (const GOp& op : G.ops)
    if (op.code == kReshape) {
      const int a = op.inputs[0], o = op.outputs[0];
      if (G.tensors[a].is_const()) return fail(TOD_ERR_UNSUPPORTED, "RESHAPE of a constant tensor");
      if (G.tensors[a].elems() * G.tensors[a].elem_size() != G.tensors[o].elems() * G.tensors[o].elem_size())
        return fail(TOD_ERR_MODEL, "RESHAPE changes the element count");
      alias_of[o] = a;
    }
  struct CopyJob { int op, src, dst; int64_t off, bytes; };
  std::vector<CopyJob> copies;
  for (size_t i = 0; i < G.ops.size(); ++i) {
    const GOp& op = G.ops[i];
    if (op.code != kConcat) continue;
    const GTensor& O = G.tensors[op.outputs[0]];
    int axis = op.axis < 0 ? op.axis + O.rank : op.axis;
    if (axis < 0 || axis >= O.rank) return fail(TOD_ERR_MODEL, "CONCATENATION axis out of range");
    const int a4 = axis + 4 - O.rank;
    int64_t outer = 1, inner = O.elem_size();
    for (int d = 0; d < a4; ++d) outer *= O.dims[d];
    for (int d = a4 + 1; d < 4; ++d) inner *= O.dims[d];
    if (outer != 1) return fail(TOD_ERR_UNSUPPORTED, "CONCATENATION along an inner axis (outer size %lld) is not supported", (long long)outer);
    int64_t off = 0;
    for (int t : op.inputs) {
      const GTensor& I = G.tensors[t];
      if (I.scale() != O.scale() || I.zp() != O.zp() || I.type != O.type)
        return fail(TOD_ERR_UNSUPPORTED, "CONCATENATION inputs must share the output quantisation");
      const int64_t chunk = int64_t(I.dims[a4]) * inner;
      const int s = storage(t);
      const bool can_place = parent[s].tensor < 0 && !G.tensors[s].is_const() && s != G.inputs[0] && s != storage(op.outputs[0]);
      if (can_place) parent[s] = Parent{op.outputs[0], off};
      else copies.push_back(CopyJob{int(i), t, op.outputs[0], off, chunk});
      off += chunk;
    }
  }

  // ---- channel padding.  A 16- or 24-channel NHWC tensor cannot feed the tensor-core path as it is: TMA needs a
  // 16-byte pixel stride, and a K chunk that is half out-of-bounds fill loads several times slower than a full one
  // (measured: 16 -> 96 at 112x112 took 133 us against 58 us for 64 -> 96).  Such tensors, when every producer is a
  // CONV_2D / ADD and every consumer a CONV_2D / ADD, are stored with 32 channels per pixel: the producing convolution
  // gets zero weight rows (the extra lanes hold the output zero point), the consuming one zero weight columns, and an
  // ADD runs over the padded buffers of all three of its tensors.  tod_yolact_fetch_tensor strips the padding.
  std::vector<int> cstore(nt, 0);
  for (int t = 0; t < nt; ++t) cstore[t] = G.tensors[t].dims[3];
  {
    std::vector<char> cand(nt, 0);
    for (int t = 0; t < nt; ++t) {
      const GTensor& X = G.tensors[t];
      if (X.is_const() || X.type != kI8 || X.dims[0] != 1 || alias_of[t] >= 0 || parent[t].tensor >= 0 || producer[t] < 0) continue;
      const int C = X.dims[3];
      if (!(C % 16 != 0 || C == 16) || C > 256) continue;
      if (std::find(G.outputs.begin(), G.outputs.end(), t) != G.outputs.end()) continue;
      const int pc = G.ops[producer[t]].code;
      cand[t] = (pc == kConv2D || pc == kAdd) ? 1 : 0;
    }
    bool changed = true;
    while (changed) {
      changed = false;
      for (const GOp& op : G.ops) {
        bool touches = false;
        for (int t : op.inputs) touches |= (t >= 0 && cand[t]);
        for (int t : op.outputs) touches |= (cand[t] != 0);
        if (!touches) continue;
        bool ok = true;
        if (op.code == kConv2D) {
          // only the activation input / output may be padded (weights and bias are constants, never candidates)
          ok = op.stride_h == 1 && op.stride_w == 1;
        } else if (op.code == kAdd) {
          ok = cand[op.inputs[0]] && cand[op.inputs[1]] && cand[op.outputs[0]] &&
               G.tensors[op.inputs[0]].dims[3] == G.tensors[op.outputs[0]].dims[3] && G.tensors[op.inputs[1]].dims[3] == G.tensors[op.outputs[0]].dims[3];
        } else {
          ok = false;
        }
        if (ok) continue;
        for (int t : op.inputs)
          if (t >= 0 && cand[t]) { cand[t] = 0; changed = true; }
        for (int t : op.outputs)
          if (cand[t]) { cand[t] = 0; changed = true; }
      }
    }
    for (int t = 0; t < nt; ++t)
      if (cand[t]) cstore[t] = int(round_up(G.tensors[t].dims[3], 32));
  }

  // ---- activation arena: one block per root storage, [max_tiles][stride]
  std::vector<int64_t> root_off(nt, -1), root_stride(nt, 0);
  int64_t total = 0;
  const int64_t mt = y->opt.max_tiles;
  for (int t = 0; t < nt; ++t) {
    const GTensor& X = G.tensors[t];
    if (X.is_const() || alias_of[t] >= 0 || parent[t].tensor >= 0) continue;
    if (producer[t] < 0 && t != G.inputs[0]) continue;  // unused
    if (X.dims[0] != 1) return fail(TOD_ERR_UNSUPPORTED, "tensor '%s' has batch %d; the model must be exported with batch 1", X.name.c_str(), X.dims[0]);
    // tile stride = dense size rounded to 16 B (what TMA needs): NHWC tensors with C % 16 == 0 stay dense, so a
    // 1x1 convolution over the whole batch is one flat GEMM
    root_stride[t] = round_up(X.elems() / std::max(1, X.dims[3]) * cstore[t] * X.elem_size(), 16);
    root_off[t] = total;
    total = round_up(total + root_stride[t] * mt, 1024);
  }
  y->act_bytes = size_t(total);
  // + 256 bytes: the stem kernel reads whole aligned words and may touch up to three bytes past a tensor's last pixel
  TOD_CUDA(cudaMalloc(&y->d_act, y->act_bytes + 256));
  y->place.assign(nt, Place{});
  for (int t = 0; t < nt; ++t) {
    const GTensor& X = G.tensors[t];
    if (X.is_const()) continue;
    int s = storage(t);
    int64_t off = 0;
    int guard = 0;
    while (parent[s].tensor >= 0 && guard++ < nt) {
      off += parent[s].off;
      s = storage(parent[s].tensor);
    }
    if (root_off[s] < 0) continue;
    y->place[t] = Place{y->d_act + root_off[s] + off, root_stride[s], X.elems() / std::max(1, X.dims[3]) * cstore[t] * X.elem_size(), X.dims[3], cstore[t]};
  }

  // ---- steps
  std::vector<bool> pad_folded(G.ops.size(), false);
  for (size_t i = 0; i < G.ops.size(); ++i) {
    const GOp& op = G.ops[i];
    Step st{};
    st.op = int(i);
    switch (op.code) {
      case kConv2D:
      case kDepthwise: {
        const bool dw = op.code == kDepthwise;
        if (op.inputs.size() < 2) return fail(TOD_ERR_MODEL, "conv operator %zu has %zu inputs", i, op.inputs.size());
        int src = op.inputs[0];
        const GTensor& Wt = G.tensors[op.inputs[1]];
        const int bias_t = op.inputs.size() > 2 ? op.inputs[2] : -1;
        const GTensor& O = G.tensors[op.outputs[0]];
        const GTensor* I = &G.tensors[src];
        if (I->type != kI8 || Wt.type != kI8 || O.type != kI8 || !Wt.is_const())
          return fail(TOD_ERR_UNSUPPORTED, "conv operator %zu: only int8 activations with constant int8 weights are supported", i);
        if (dw && op.depth_multiplier != 1) return fail(TOD_ERR_UNSUPPORTED, "depthwise multiplier %d", op.depth_multiplier);
        for (int64_t z : Wt.zero_points)
          if (z != 0) return fail(TOD_ERR_UNSUPPORTED, "conv operator %zu: weight zero point must be 0", i);
        ConvGeom g{};
        g.KH = Wt.dims[1];
        g.KW = Wt.dims[2];
        g.OC = dw ? Wt.dims[3] : Wt.dims[0];
        g.stride_h = op.stride_h; g.stride_w = op.stride_w; g.dil_h = op.dil_h; g.dil_w = op.dil_w;
        int OH, OW, ph, pw;
        conv_out_pad(op.padding, I->dims[1], g.KH, g.stride_h, g.dil_h, &OH, &ph);
        conv_out_pad(op.padding, I->dims[2], g.KW, g.stride_w, g.dil_w, &OW, &pw);
        if (O.dims[1] != OH || O.dims[2] != OW || O.dims[3] != g.OC || (!dw && Wt.dims[3] != I->dims[3]) || (dw && Wt.dims[3] != I->dims[3]))
          return fail(TOD_ERR_MODEL, "conv operator %zu ('%s'): inconsistent shapes", i, O.name.c_str());
        g.pad_top = ph;
        g.pad_left = pw;
        // PAD folding: a zero-point PAD feeding only this VALID convolution contributes nothing to the
        // accumulators (in - zp == 0), which is exactly "tap skipped" on the unpadded tensor.
        const int pp = producer[src];
        if (fuse && op.padding == kValid && pp >= 0 && G.ops[pp].code == kPad && consumers[src] == 1) {
          const GOp& P = G.ops[pp];
          const GTensor& PI = G.tensors[P.inputs[0]];
          const GTensor& PD = G.tensors[P.inputs[1]];
          if (PD.is_const() && PD.type == kI32 && PD.elems() == 8 && PI.scale() == I->scale() && PI.zp() == I->zp()) {
            int32_t pd[8];
            std::memcpy(pd, PD.const_data, 32);
            if (!pd[0] && !pd[1] && !pd[6] && !pd[7]) {
              g.pad_top = pd[2];
              g.pad_left = pd[4];
              src = P.inputs[0];
              I = &G.tensors[src];
              pad_folded[pp] = true;
            }
          }
        }
        g.IH = I->dims[1]; g.IW = I->dims[2]; g.IC = I->dims[3];
        g.OH = OH; g.OW = OW;
        st.g = g;
        st.in0 = src;
        st.out = op.outputs[0];
        st.in_zp = I->zp();
        st.out_zp = O.zp();
        activation_range(op.activation, O, &st.act_min, &st.act_max);
        const int ICr = g.IC, OCr = g.OC;                       // real channel counts
        const int ICs = dw ? ICr : cstore[src], OCs = dw ? OCr : cstore[op.outputs[0]];  // stored (padded) ones
        const int taps = g.KH * g.KW;
        {
          // requantisation tables; padding lanes copy channel 0's constants so a layer stays eligible for the fast epilogue
          std::vector<int32_t> q(OCs), sh(OCs);
          for (int c = 0; c < OCs; ++c) {
            const int cr = c < OCr ? c : 0;
            const float ws = Wt.scales.size() > 1 ? Wt.scales[cr] : Wt.scale();
            const double eff = double(I->scale()) * double(ws) / double(O.scale());
            int sft = 0;
            quantize_multiplier(eff, &q[c], &sft);
            sh[c] = sft;
          }
          st.mult_off = arena->add(q.data(), q.size() * 4);
          st.shift_off = arena->add(sh.data(), sh.size() * 4);
          // the 4-instruction requantisation (Requant::fast_tab) when every channel qualifies
          bool fast_ok = std::abs(O.zp()) <= 128;
          for (int c = 0; c < OCs && fast_ok; ++c) fast_ok = sh[c] <= -1 && sh[c] >= -22 && q[c] >= 0;
          if (fast_ok) {
            std::vector<int32_t> ft(size_t(OCs) * 4);
            for (int c = 0; c < OCs; ++c) {
              const int rs = -sh[c];
              ft[size_t(c) * 4 + 0] = q[c];
              ft[size_t(c) * 4 + 1] = rs;
              ft[size_t(c) * 4 + 2] = int32_t(0x80000000u);
              ft[size_t(c) * 4 + 3] = (1 << (rs - 1)) + O.zp() * (1 << rs);
            }
            st.fast_off = arena->add(ft.data(), ft.size() * 4);
          }
        }
        const int8_t* w_real = reinterpret_cast<const int8_t*>(Wt.const_data);
        std::vector<int8_t> w_pad;
        if (ICs != ICr || OCs != OCr) {  // OHWI with zero columns / rows for the padding lanes
          w_pad.assign(size_t(OCs) * taps * ICs, 0);
          for (int oc = 0; oc < OCr; ++oc)
            for (int tp = 0; tp < taps; ++tp)
              std::memcpy(&w_pad[(size_t(oc) * taps + tp) * ICs], w_real + (size_t(oc) * taps + tp) * ICr, size_t(ICr));
          st.w_off = arena->add(w_pad.data(), w_pad.size());
        } else {
          st.w_off = arena->add(Wt.const_data, Wt.const_bytes);
        }
        if (bias_t >= 0 || OCs != OCr) {
          std::vector<int32_t> b(OCs, 0);
          if (bias_t >= 0) {
            const GTensor& B = G.tensors[bias_t];
            if (!B.is_const() || B.type != kI32 || B.elems() != OCr) return fail(TOD_ERR_UNSUPPORTED, "conv operator %zu: bias must be constant int32[OC]", i);
            std::memcpy(b.data(), B.const_data, size_t(OCr) * 4);
          }
          st.bias_off = arena->add(b.data(), b.size() * 4);
        }
        if (dw) {
          st.kind = kStepDepthwise;
          st.macs = int64_t(OH) * OW * g.OC * g.KH * g.KW;
        } else {
          st.kind = kStepConvDirect;
          std::vector<int32_t> wsum(size_t(OCs) * taps, 0);
          for (int oc = 0; oc < OCr; ++oc)
            for (int tp = 0; tp < taps; ++tp) {
              int32_t sm = 0;
              for (int ic = 0; ic < ICr; ++ic) sm += w_real[(size_t(oc) * taps + tp) * ICr + ic];
              wsum[size_t(oc) * taps + tp] = sm;
            }
          st.wsum_off = arena->add(wsum.data(), wsum.size() * 4);
          st.macs = int64_t(OH) * OW * OCr * taps * ICr;
          st.g.IC = ICs;
          st.g.OC = OCs;
        }
        break;
      }
      case kAdd: {
        const GTensor& A = G.tensors[op.inputs[0]];
        const GTensor& B = G.tensors[op.inputs[1]];
        const GTensor& O = G.tensors[op.outputs[0]];
        if (A.elems() != B.elems() || A.elems() != O.elems() || A.type != kI8 || B.type != kI8 || O.type != kI8 || A.is_const() || B.is_const())
          return fail(TOD_ERR_UNSUPPORTED, "ADD operator %zu: only same-shape int8 activations are supported", i);
        st.kind = kStepAdd;
        st.in0 = op.inputs[0]; st.in1 = op.inputs[1]; st.out = op.outputs[0];
        const double twice_max = 2 * std::max(A.scale(), B.scale());  // tensorflow/lite/kernels/add.cc Prepare
        int s;
        quantize_multiplier(A.scale() / twice_max, &st.add.mult_a, &s); st.add.shift_a = s;
        quantize_multiplier(B.scale() / twice_max, &st.add.mult_b, &s); st.add.shift_b = s;
        quantize_multiplier(twice_max / ((1 << 20) * O.scale()), &st.add.mult_out, &s); st.add.shift_out = s;
        st.add.zp_a = A.zp(); st.add.zp_b = B.zp(); st.add.zp_out = O.zp();
        activation_range(op.activation, O, &st.add.act_min, &st.add.act_max);
        break;
      }
      case kQuantize:
      case kRelu:
      case kTanh: {
        const GTensor& I = G.tensors[op.inputs[0]];
        const GTensor& O = G.tensors[op.outputs[0]];
        const bool in8 = I.type == kI8 || I.type == kU8, out8 = O.type == kI8 || O.type == kU8;
        if (!in8 || !out8 || I.is_const() || I.elems() != O.elems())
          return fail(TOD_ERR_UNSUPPORTED, "operator %zu (code %d): only 8-bit -> 8-bit activations are supported", i, op.code);
        uint8_t lut[256];
        const int32_t imin = I.type == kU8 ? 0 : -128;
        const int32_t omin = O.type == kU8 ? 0 : -128, omax = O.type == kU8 ? 255 : 127;
        if (op.code == kTanh) {
          if (I.type != O.type) return fail(TOD_ERR_UNSUPPORTED, "TANH operator %zu changes the tensor type", i);
          const float inv = 1 / O.scale();  // activations.cc PopulateLookupTable
          for (int k = 0; k < 256; ++k) {
            const int32_t v = imin + k;
            const float deq = I.scale() * (v - I.zp());
            const float tr = std::tanh(deq);
            const float resc = std::round(tr * inv);
            const int32_t qv = int32_t(resc + O.zp());
            lut[uint8_t(v)] = uint8_t(std::max(std::min(omax, qv), omin));
          }
        } else {
          int32_t q; int sh;
          quantize_multiplier(double(I.scale()) / double(O.scale()), &q, &sh);
          int32_t lo = omin;
          if (op.code == kRelu) lo = std::max(omin, O.zp() + int32_t(std::round(0.0f / O.scale())));
          for (int k = 0; k < 256; ++k) {
            const int32_t v = imin + k;
            int32_t r = mul_by_quant_mult(v - I.zp(), q, sh) + O.zp();
            r = std::min(std::max(r, lo), omax);
            lut[uint8_t(v)] = uint8_t(r);
          }
        }
        st.kind = kStepLut;
        st.in0 = op.inputs[0]; st.out = op.outputs[0];
        st.lut_off = arena->add(lut, 256);
        st.lut_host.assign(lut, lut + 256);
        break;
      }
      case kPad: {
        const GTensor& I = G.tensors[op.inputs[0]];
        const GTensor& PD = G.tensors[op.inputs[1]];
        const GTensor& O = G.tensors[op.outputs[0]];
        if (I.type != kI8 || !PD.is_const() || PD.type != kI32 || PD.elems() != 8) return fail(TOD_ERR_UNSUPPORTED, "PAD operator %zu: need int8 input and constant int32[4,2] paddings", i);
        int32_t pd[8];
        std::memcpy(pd, PD.const_data, 32);
        if (pd[0] || pd[1] || pd[6] || pd[7]) return fail(TOD_ERR_UNSUPPORTED, "PAD operator %zu: batch / channel padding", i);
        if (O.dims[1] != I.dims[1] + pd[2] + pd[3] || O.dims[2] != I.dims[2] + pd[4] + pd[5] || O.dims[3] != I.dims[3])
          return fail(TOD_ERR_MODEL, "PAD operator %zu: inconsistent shapes", i);
        st.kind = kStepPad;
        st.in0 = op.inputs[0]; st.out = op.outputs[0];
        st.pad_top = pd[2]; st.pad_left = pd[4];
        st.fill = int8_t(O.zp());
        break;
      }
      case kResizeBilinear: {
        const GTensor& I = G.tensors[op.inputs[0]];
        const GTensor& O = G.tensors[op.outputs[0]];
        if (I.type != kI8 || O.type != kI8 || I.dims[3] != O.dims[3] || I.dims[3] % 4 != 0)
          return fail(TOD_ERR_UNSUPPORTED, "RESIZE_BILINEAR operator %zu: need int8 with channels %% 4 == 0", i);
        st.kind = kStepResize;
        st.in0 = op.inputs[0]; st.out = op.outputs[0];
        st.align_corners = op.align_corners; st.half_pixel = op.half_pixel_centers;
        break;
      }
      case kReshape:
      case kConcat:
        continue;  // aliases; residual copies are appended below
      default:
        return fail(TOD_ERR_UNSUPPORTED, "operator code %d", op.code);
    }
    y->steps.push_back(st);
    // copies for concat inputs that could not be placed run right after their concat's position
    (void)copies;
  }
  // un-aliased concat inputs: a copy step after the producer has run (append in op order)
  for (const CopyJob& c : copies) {
    Step st{};
    st.kind = kStepCopy;
    st.op = c.op;
    st.in0 = c.src; st.out = c.dst;
    st.copy_dst_off = c.off; st.copy_bytes = c.bytes;
    // insert after the last step whose op index < concat op index
    auto it = std::find_if(y->steps.begin(), y->steps.end(), [&](const Step& s) { return s.op > c.op; });
    y->steps.insert(it, st);
  }
  // drop folded PADs
  y->steps.erase(std::remove_if(y->steps.begin(), y->steps.end(), [&](const Step& s) { return s.kind == kStepPad && pad_folded[s.op]; }),
                 y->steps.end());
  for (const Step& s : y->steps) {
    y->macs_per_tile += s.macs;
    if (s.in0 >= 0 && !y->place[s.in0].base) return fail(TOD_ERR_MODEL, "operator %d reads tensor %d which nothing produces", s.op, s.in0);
  }
  // ---- byte-map fusion: QUANTIZE / RELU / TANH are pure functions of one byte.  A chain  conv -> [RESHAPE | placed
  // CONCATENATION | byte map]*  whose tensors have no other reader collapses into the conv's epilogue: the maps are
  // composed on the host into one 256-entry table and the conv writes straight into the chain's last buffer (e.g. the
  // uint8 class / box / coefficient outputs come directly from the five head levels' convolutions).
  y->fused_away.assign(nt, 0);
  if (fuse) {
    std::vector<int> op_of_tensor_consumer_count(consumers);  // readers per tensor (graph outputs count as one)
    auto single_reader = [&](int t) { return op_of_tensor_consumer_count[t] == 1; };
    bool changed = true;
    while (changed) {
      changed = false;
      for (size_t li = 0; li < y->steps.size(); ++li) {
        Step& L = y->steps[li];
        if (L.kind != kStepLut) continue;
        // walk back from the map's input through aliases to the producing steps
        std::vector<int> srcs;  // conv step indices
        bool ok = true;
        std::function<void(int)> back = [&](int t) {
          if (!ok) return;
          if (!single_reader(t) || G.tensors[t].is_const() || t == G.inputs[0]) {
            ok = false;
            return;
          }
          int found = -1;
          for (size_t k = 0; k < y->steps.size(); ++k)
            if (y->steps[k].out == t && y->steps[k].kind != kStepCopy) found = int(k);
          if (found >= 0) {  // a kernel writes t (possibly a conv that already absorbed an earlier map)
            const StepKind kd = y->steps[found].kind;
            if (kd != kStepConvDirect && kd != kStepDepthwise) ok = false;
            else srcs.push_back(found);
            return;
          }
          const int po = producer[t];
          if (po < 0) {
            ok = false;
            return;
          }
          const GOp& op = G.ops[po];
          if (op.code == kReshape) return back(op.inputs[0]);
          if (op.code == kConcat) {
            for (const CopyJob& c : copies)
              if (c.op == po) ok = false;  // an input that had to be copied: keep the concat buffer
            for (int in : op.inputs) back(in);
            return;
          }
          ok = false;
        };
        back(L.in0);
        if (!ok || srcs.empty()) continue;
        const Place& lin = y->place[L.in0];
        const Place& lout = y->place[L.out];
        for (int si : srcs) {
          Step& P = y->steps[si];
          const Place cur = P.out_moved ? P.out_place : y->place[P.out];
          Place np;
          np.base = lout.base + (cur.base - lin.base);
          np.tile_stride = lout.tile_stride;
          np.bytes = cur.bytes;
          std::vector<uint8_t> composed(256);
          for (int b = 0; b < 256; ++b) composed[b] = L.lut_host[P.lut_host.empty() ? b : P.lut_host[b]];
          P.lut_host = composed;
          y->fused_away[P.out] = 1;
          P.out_moved = true;
          P.out_place = np;
          P.out = L.out;  // for dependency tracking and fetches the conv now *is* the producer of the map's output
        }
        // tensors on the way are no longer written
        std::function<void(int)> mark = [&](int t) {
          y->fused_away[t] = 1;
          const int po = producer[t];
          if (po < 0) return;
          const GOp& op = G.ops[po];
          if (op.code == kReshape) mark(op.inputs[0]);
          else if (op.code == kConcat)
            for (int in : op.inputs) mark(in);
        };
        mark(L.in0);
        y->steps.erase(y->steps.begin() + li);
        changed = true;
        break;
      }
    }
  }
  // ---- input QUANTIZE folding: the graph starts with uint8 -> int8 (x - 128, same scale), i.e. the byte map x ^ 0x80.  When
  // that tensor's only reader is the RGB stem kernel, the stem reads the uint8 tiles itself and flips the bit on the fly.
  if (fuse && y->opt.conv_impl != 2) {  // conv_impl 2 runs without Requant::fast_tab, which the stem kernel needs
    for (size_t qi = 0; qi < y->steps.size(); ++qi) {
      const Step& Q = y->steps[qi];
      if (Q.kind != kStepLut || Q.in0 != G.inputs[0] || consumers[Q.out] != 1) continue;
      bool is_flip = true;
      for (int b = 0; b < 256 && is_flip; ++b) is_flip = Q.lut_host[b] == uint8_t(b ^ 0x80);
      if (!is_flip) continue;
      for (size_t si = 0; si < y->steps.size(); ++si) {
        Step& S = y->steps[si];
        if (S.kind != kStepConvDirect || S.in0 != Q.out || !S.lut_host.empty() || S.fast_off < 0 || S.fused_add) continue;
        const Place& pin = y->place[Q.in0];
        const Place& pout = S.out_moved ? S.out_place : y->place[S.out];
        Requant probe{nullptr, nullptr, S.out_zp, S.act_min, S.act_max, nullptr};
        probe.fast_tab = reinterpret_cast<const int4*>(uintptr_t(16));  // eligibility only looks at presence
        if (!stem_kernel_eligible(S.g, probe, pin.base, pin.tile_stride, pout.base, pout.tile_stride)) continue;
        S.in0 = Q.in0;
        S.g.in_xor = 0x80;
        y->fused_away[Q.out] = 1;
        y->steps.erase(y->steps.begin() + qi);
        qi = y->steps.size();  // one input, one stem
        break;
      }
    }
  }
  // ---- residual / FPN ADD fusion: an ADD one of whose inputs is a tensor-core convolution's only-read output runs in
  // that convolution's epilogue (two 256-entry rescale tables + one output rescale, the arithmetic of ops.cu::add_kernel).
  // The convolution moves to the ADD's position in the step list, so the other operand is always produced before it.
  if (fuse && y->opt.conv_impl == 0) {
    for (size_t ai = 0; ai < y->steps.size(); ++ai) {
      if (y->steps[ai].kind != kStepAdd) continue;
      const Step A = y->steps[ai];
      int pick = -1;
      bool conv_is_a = true;
      for (int side = 0; side < 2 && pick < 0; ++side) {
        const int t = side == 0 ? A.in0 : A.in1;
        if (consumers[t] != 1) continue;
        for (size_t k = 0; k < ai; ++k) {
          const Step& P = y->steps[k];
          if (P.kind != kStepConvDirect || P.out != t || P.out_moved || P.fused_add || !P.lut_host.empty() || P.fast_off < 0) continue;
          const Place& pin = y->place[P.in0];
          const Place& pout = y->place[P.out];
          const Place& padd = y->place[A.out];
          const Place& pres = y->place[side == 0 ? A.in1 : A.in0];
          if (!conv_tc_supported(P.g, pin.tile_stride, pin.base, reinterpret_cast<const void*>(uintptr_t(256)))) continue;
          if (P.g.OC % 16 != 0 || (reinterpret_cast<uintptr_t>(padd.base) & 15) || (padd.tile_stride & 15) || (reinterpret_cast<uintptr_t>(pres.base) & 15) ||
              (pres.tile_stride & 15))
            continue;
          if (padd.bytes != pout.bytes || pres.bytes != pout.bytes || padd.c_store != pout.c_store || pres.c_store != pout.c_store) continue;
          pick = int(k);
          conv_is_a = side == 0;
        }
      }
      if (pick < 0) continue;
      Step P = y->steps[pick];
      P.fused_add = true;
      P.add_conv_is_a = conv_is_a;
      P.add = A.add;
      P.in1 = conv_is_a ? A.in1 : A.in0;
      y->fused_away[P.out] = 1;
      P.out_moved = true;
      P.out_place = y->place[A.out];
      P.out = A.out;
      y->steps[ai] = P;                                 // the conv takes the ADD's slot ...
      y->steps.erase(y->steps.begin() + pick);          // ... and leaves its own (pick < ai)
      --ai;
    }
  }
  // ---- sibling-head fusion: two convolutions over the SAME input with the same geometry, one of them at most 16 channels wide
  // (the 12-column box head beside the 96-column coefficient head of every pyramid level).  The narrow one's time is all
  // A-operand traffic - 3x3 x 256 channels re-read per tap for a 16-column MMA - so its weights become one more 16-column chunk
  // of the wide one's accumulator: one launch, one pass over the activations, and the narrow layer's bytes leave from the
  // epilogue registers into its own output (conv_tc.cu, TcParams::x_*).  Needs the tcgen05 fast epilogue for both (checked
  // here with the planner's own criteria) and 16-byte rows for the wide one.
  if (fuse && y->opt.conv_impl == 0 && !(std::getenv("TOD_HEAD_MERGE") && std::atoi(std::getenv("TOD_HEAD_MERGE")) == 0)) {
    auto full_range = [](const Step& s) { return s.act_min == -128 && s.act_max == 127; };
    for (size_t hi = 0; hi < y->steps.size(); ++hi) {
      Step& H = y->steps[hi];
      if (H.kind != kStepConvDirect || H.fused_add || H.fast_off < 0 || H.sib_cols || !full_range(H)) continue;
      if (H.g.OC % 16 != 0 || H.g.OC + 16 > 256 || H.g.OC < 32) continue;
      const Place hp = H.out_moved ? H.out_place : y->place[H.out];
      if ((reinterpret_cast<uintptr_t>(hp.base) & 15) || (hp.tile_stride & 15) || (!H.out_moved && hp.c_store != H.g.OC)) continue;
      const Place& hin = y->place[H.in0];
      if (!conv_tc_supported(H.g, hin.tile_stride, hin.base, reinterpret_cast<const void*>(uintptr_t(256)))) continue;
      if (H.g.KH * H.g.KW == 1 && hin.tile_stride == int64_t(H.g.IH) * H.g.IW * H.g.IC) continue;  // flat 1x1 layers take the flat-linear kernels
      int pick = -1;
      for (size_t bi = 0; bi < y->steps.size() && pick < 0; ++bi) {
        const Step& B = y->steps[bi];
        if (bi == hi || B.kind != kStepConvDirect || B.in0 != H.in0 || B.fused_add || B.fast_off < 0 || B.sib_cols || !full_range(B)) continue;
        if (B.g.OC > 16 || B.g.OC % 4 != 0 || B.g.OC < 4) continue;
        const ConvGeom &a = H.g, &b = B.g;
        if (a.KH != b.KH || a.KW != b.KW || a.stride_h != b.stride_h || a.stride_w != b.stride_w || a.dil_h != b.dil_h || a.dil_w != b.dil_w ||
            a.pad_top != b.pad_top || a.pad_left != b.pad_left || a.IH != b.IH || a.IW != b.IW || a.IC != b.IC || a.OH != b.OH || a.OW != b.OW || a.in_xor != b.in_xor)
          continue;
        const Place bp = B.out_moved ? B.out_place : y->place[B.out];
        if ((reinterpret_cast<uintptr_t>(bp.base) & 3) || (bp.tile_stride & 3) || (!B.out_moved && bp.c_store != B.g.OC)) continue;
        pick = int(bi);
      }
      if (pick < 0) continue;
      const Step B = y->steps[pick];
      const int OCh = H.g.OC, OCb = B.g.OC, OCt = OCh + 16, taps = H.g.KH * H.g.KW, IC = H.g.IC;
      auto host = [&](int64_t off) { return arena->host.data() + off; };
      // concatenated constants: host rows, sibling rows, zero rows (padding lanes copy the sibling's channel 0 requantisation)
      std::vector<int8_t> w(size_t(OCt) * taps * IC, 0);
      std::memcpy(w.data(), host(H.w_off), size_t(OCh) * taps * IC);
      std::memcpy(w.data() + size_t(OCh) * taps * IC, host(B.w_off), size_t(OCb) * taps * IC);
      std::vector<int32_t> bias(OCt, 0), wsum(size_t(OCt) * taps, 0), mult(OCt), shift(OCt);
      if (H.bias_off >= 0) std::memcpy(bias.data(), host(H.bias_off), size_t(OCh) * 4);
      if (B.bias_off >= 0) std::memcpy(bias.data() + OCh, host(B.bias_off), size_t(OCb) * 4);
      std::memcpy(wsum.data(), host(H.wsum_off), size_t(OCh) * taps * 4);
      std::memcpy(wsum.data() + size_t(OCh) * taps, host(B.wsum_off), size_t(OCb) * taps * 4);
      const int32_t *hm = reinterpret_cast<const int32_t*>(host(H.mult_off)), *hs = reinterpret_cast<const int32_t*>(host(H.shift_off));
      const int32_t *bm = reinterpret_cast<const int32_t*>(host(B.mult_off)), *bs = reinterpret_cast<const int32_t*>(host(B.shift_off));
      for (int c = 0; c < OCt; ++c) {
        mult[c] = c < OCh ? hm[c] : bm[c - OCh < OCb ? c - OCh : 0];
        shift[c] = c < OCh ? hs[c] : bs[c - OCh < OCb ? c - OCh : 0];
      }
      Step M = H;
      M.w_off = arena->add(w.data(), w.size());
      M.bias_off = arena->add(bias.data(), bias.size() * 4);
      M.wsum_off = arena->add(wsum.data(), wsum.size() * 4);
      M.mult_off = arena->add(mult.data(), mult.size() * 4);
      M.shift_off = arena->add(shift.data(), shift.size() * 4);
      M.g.OC = OCt;
      M.macs = H.macs + B.macs;
      M.sib_out = B.out;
      M.sib_cols = OCb;
      M.sib_host_oc = OCh;
      M.sib_zp = B.out_zp;
      M.sib_place = B.out_moved ? B.out_place : y->place[B.out];
      M.sib_lut_host = B.lut_host;
      if (M.lut_host.empty() && !B.lut_host.empty()) {   // the epilogue applies byte maps per chunk: the host layer then needs one too
        M.lut_host.resize(256);
        for (int b = 0; b < 256; ++b) M.lut_host[b] = uint8_t(b);
      }
      y->steps[hi] = M;
      y->steps.erase(y->steps.begin() + pick);
      if (size_t(pick) < hi) --hi;
    }
  }
  for (Step& st : y->steps) {
    if ((st.kind == kStepConvDirect || st.kind == kStepDepthwise) && !st.lut_host.empty()) st.post_lut_off = arena->add(st.lut_host.data(), 256);
    if (st.sib_cols && !st.sib_lut_host.empty()) st.sib_lut_off = arena->add(st.sib_lut_host.data(), 256);
  }
  // ---- ReLU-type layers (activation floor at or above the output zero point, no byte map / ADD behind them): the CUDA-core
  // kernels that read Requant::fast_tab (depthwise, RGB stem) get the two-instruction form (fixedpoint.cuh::requant_relu)
  if (y->opt.conv_impl != 2 && !(std::getenv("TOD_RELU_TAB") && std::atoi(std::getenv("TOD_RELU_TAB")) == 0))
    for (Step& st : y->steps) {
      if ((st.kind != kStepConvDirect && st.kind != kStepDepthwise) || st.fast_off < 0 || !st.lut_host.empty() || st.fused_add) continue;
      if (st.act_min < st.out_zp) continue;
      int32_t* ft = reinterpret_cast<int32_t*>(arena->host.data() + st.fast_off);
      const int n = st.kind == kStepDepthwise ? st.g.OC : st.g.OC;
      for (int c = 0; c < n; ++c) {
        const int32_t q = ft[4 * c], rs = ft[4 * c + 1];
        const int64_t A = relu_addend(q, rs, st.out_zp);
        ft[4 * c + 1] = rs - 1;
        ft[4 * c + 2] = int32_t(uint32_t(uint64_t(A) & 0xFFFFFFFFull));
        ft[4 * c + 3] = int32_t(uint32_t(uint64_t(A) >> 32));
      }
      st.relu_tab = true;
    }
  // ---- data dependencies between steps
  std::vector<std::vector<int>> op_steps(G.ops.size());  // steps generated for an op (its kernel, or a concat's copies)
  for (size_t i = 0; i < y->steps.size(); ++i) op_steps[y->steps[i].op].push_back(int(i));
  std::vector<std::vector<int>> writers(nt);  // steps whose (possibly moved) output is tensor t
  for (size_t i = 0; i < y->steps.size(); ++i) {
    if (y->steps[i].out_moved) writers[y->steps[i].out].push_back(int(i));
    if (y->steps[i].sib_cols) writers[y->steps[i].sib_out].push_back(int(i));   // the merged launch is the sibling output's producer
  }
  std::vector<std::vector<int>> memo(nt);
  std::vector<char> done(nt, 0);
  std::function<const std::vector<int>&(int)> producers = [&](int t) -> const std::vector<int>& {
    if (done[t]) return memo[t];
    done[t] = 1;
    std::vector<int> out = writers[t];
    const int po = producer[t];
    if (po >= 0 && writers[t].empty()) {
      const GOp& op = G.ops[po];
      if (op.code == kReshape || (op.code == kPad && pad_folded[po])) {
        out = producers(op.inputs[0]);
      } else if (op.code == kConcat) {
        for (int in : op.inputs) {
          const std::vector<int>& p = producers(in);
          out.insert(out.end(), p.begin(), p.end());
        }
        out.insert(out.end(), op_steps[po].begin(), op_steps[po].end());
      } else {
        out = op_steps[po];
      }
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    memo[t] = out;
    return memo[t];
  };
  for (size_t i = 0; i < y->steps.size(); ++i) {
    Step& st = y->steps[i];
    std::vector<int> d;
    for (int t : {st.in0, st.in1})
      if (t >= 0) {
        const std::vector<int>& p = producers(t);
        d.insert(d.end(), p.begin(), p.end());
      }
    if (st.kind == kStepCopy) {
      // a residual concat copy also has to follow whatever else writes the destination's other regions? No: regions
      // are disjoint; it only reads its source.
    }
    std::sort(d.begin(), d.end());
    d.erase(std::unique(d.begin(), d.end()), d.end());
    d.erase(std::remove_if(d.begin(), d.end(), [&](int v) { return v >= int(i); }), d.end());
    st.deps = d;
  }
  y->out_steps.clear();
  for (int t : G.outputs) y->out_steps.push_back(producers(t));
  return TOD_OK;
}

int upload_consts_and_bind(tod_yolact* y, const ConstArena& arena) {
  TOD_CUDA(cudaMalloc(&y->d_const, arena.host.size() + 256));
  TOD_CUDA(cudaMemcpy(y->d_const, arena.host.data(), arena.host.size(), cudaMemcpyHostToDevice));
  for (Step& st : y->steps) {
    if (st.kind != kStepConvDirect || y->opt.conv_impl == 1) continue;  // conv_impl: 0 = tcgen05, 1 = CUDA cores only, 2 = tcgen05 with the general epilogue
    const Place& pi = y->place[st.in0];
    const Place& po = st.out_moved ? st.out_place : y->place[st.out];
    const int8_t* w = reinterpret_cast<const int8_t*>(y->d_const + st.w_off);
    if (!conv_tc_supported(st.g, pi.tile_stride, pi.base, w)) {
      if (st.fused_add) return fail(TOD_ERR_UNSUPPORTED, "planner fused an ADD into a convolution the tensor-core path rejects");
      continue;
    }
    ConvTcArgs a{};
    a.min_rounds = y->opt.batches_in_flight >= 2 ? 2 : 1;
    a.g = st.g;
    a.in = reinterpret_cast<const int8_t*>(pi.base);
    a.in_tile_stride = pi.tile_stride;
    a.w = w;
    a.in_zp = st.in_zp;
    a.rq = Requant{reinterpret_cast<const int32_t*>(y->d_const + st.mult_off), reinterpret_cast<const int32_t*>(y->d_const + st.shift_off),
                   st.out_zp, st.act_min, st.act_max, st.post_lut_off >= 0 ? y->d_const + st.post_lut_off : nullptr};
    a.out = reinterpret_cast<int8_t*>(po.base);
    a.out_tile_stride = po.tile_stride;
    a.max_tiles = y->opt.max_tiles;
    a.h_bias = st.bias_off >= 0 ? reinterpret_cast<const int32_t*>(arena.host.data() + st.bias_off) : nullptr;
    a.h_wsum = reinterpret_cast<const int32_t*>(arena.host.data() + st.wsum_off);
    a.h_mult = reinterpret_cast<const int32_t*>(arena.host.data() + st.mult_off);
    a.h_shift = reinterpret_cast<const int32_t*>(arena.host.data() + st.shift_off);
    a.fast_epilogue = y->opt.conv_impl == 2 ? 0 : 1;
    if (st.sib_cols) {
      a.out_oc = st.sib_host_oc;
      a.x_cols = st.sib_cols;
      a.x_out = reinterpret_cast<int8_t*>(st.sib_place.base);
      a.x_out_tile_stride = st.sib_place.tile_stride;
      a.x_out_zp = st.sib_zp;
      a.x_lut = st.sib_lut_off >= 0 ? y->d_const + st.sib_lut_off : nullptr;
    }
    ConvTcAdd fa{};
    if (st.fused_add) {
      const Place& pr = y->place[st.in1];
      fa.resid = reinterpret_cast<const int8_t*>(pr.base);
      fa.resid_tile_stride = pr.tile_stride;
      const AddParams& ap = st.add;
      for (int b = 0; b < 256; ++b) {  // index = the int8 byte pattern; same literal arithmetic as ops.cu::add_kernel
        const int v = int(int8_t(b));
        const int32_t ta = mul_by_quant_mult((v - ap.zp_a) * (1 << 20), ap.mult_a, ap.shift_a);
        const int32_t tb = mul_by_quant_mult((v - ap.zp_b) * (1 << 20), ap.mult_b, ap.shift_b);
        fa.tab[b] = st.add_conv_is_a ? ta : tb;
        fa.tab[256 + b] = st.add_conv_is_a ? tb : ta;
      }
      fa.mult_out = ap.mult_out;
      fa.shift_out = ap.shift_out;
      fa.zp_out = ap.zp_out;
      fa.act_min = ap.act_min;
      fa.act_max = ap.act_max;
      a.add = &fa;
    }
    TOD_TRY(conv_tc_create(a, &st.tc));
    st.kind = kStepConvTc;
    y->tc_layers++;
  }
  return TOD_OK;
}

int run_step(tod_yolact* y, const Step& st, int n, cudaStream_t s) {
  const Place& pi = y->place[st.in0];
  const Place& po = st.out_moved ? st.out_place : y->place[st.out];
  // TOD_DIAG_SKIP (timing attribution only, results are garbage): 1 = no depthwise, 2 = no resize, 4 = no tcgen05 convs, 8 = no detection tail,
  // 16 = no stem / direct conv
  const int diag_skip = y->diag_skip;  // read from the environment when the handle is created
  if (diag_skip) {
    if ((diag_skip & 1) && st.kind == kStepDepthwise) return TOD_OK;
    if ((diag_skip & 2) && st.kind == kStepResize) return TOD_OK;
    if ((diag_skip & 4) && st.kind == kStepConvTc) return TOD_OK;
    if ((diag_skip & 16) && st.kind == kStepConvDirect) return TOD_OK;
  }
  switch (st.kind) {
    case kStepConvTc:
      return conv_tc_launch(st.tc, n, s);
    case kStepConvDirect:
    case kStepDepthwise: {
      Requant rq{reinterpret_cast<const int32_t*>(y->d_const + st.mult_off), reinterpret_cast<const int32_t*>(y->d_const + st.shift_off),
                 st.out_zp, st.act_min, st.act_max, st.post_lut_off >= 0 ? y->d_const + st.post_lut_off : nullptr};
      if (st.fast_off >= 0 && y->opt.conv_impl != 2) {
        rq.fast_tab = reinterpret_cast<const int4*>(y->d_const + st.fast_off);
        rq.relu_tab = st.relu_tab;
      }
      const int8_t* w = reinterpret_cast<const int8_t*>(y->d_const + st.w_off);
      const int32_t* bias = st.bias_off >= 0 ? reinterpret_cast<const int32_t*>(y->d_const + st.bias_off) : nullptr;
      if (st.kind == kStepDepthwise)
        launch_depthwise(reinterpret_cast<const int8_t*>(pi.base), pi.tile_stride, w, bias, st.in_zp, st.g, rq,
                         reinterpret_cast<int8_t*>(po.base), po.tile_stride, n, s);
      else
        launch_conv_direct(reinterpret_cast<const int8_t*>(pi.base), pi.tile_stride, w, bias,
                           reinterpret_cast<const int32_t*>(y->d_const + st.wsum_off), st.in_zp, st.g, rq,
                           reinterpret_cast<int8_t*>(po.base), po.tile_stride, n, s);
      break;
    }
    case kStepAdd: {
      const Place& pb = y->place[st.in1];
      launch_add(reinterpret_cast<const int8_t*>(pi.base), pi.tile_stride, reinterpret_cast<const int8_t*>(pb.base), pb.tile_stride,
                 reinterpret_cast<int8_t*>(po.base), po.tile_stride, po.bytes, n, st.add, s);
      break;
    }
    case kStepLut:
      launch_lut(pi.base, pi.tile_stride, po.base, po.tile_stride, po.bytes, n, y->d_const + st.lut_off, s);
      break;
    case kStepPad: {
      const GTensor& I = y->T(st.in0);
      const GTensor& O = y->T(st.out);
      launch_pad(reinterpret_cast<const int8_t*>(pi.base), pi.tile_stride, I.dims[1], I.dims[2], I.dims[3], st.pad_top, st.pad_left,
                 O.dims[1], O.dims[2], st.fill, reinterpret_cast<int8_t*>(po.base), po.tile_stride, n, s);
      break;
    }
    case kStepResize: {
      const GTensor& I = y->T(st.in0);
      const GTensor& O = y->T(st.out);
      launch_resize_bilinear(reinterpret_cast<const int8_t*>(pi.base), pi.tile_stride, I.dims[1], I.dims[2], I.dims[3],
                             reinterpret_cast<int8_t*>(po.base), po.tile_stride, O.dims[1], O.dims[2], st.align_corners, st.half_pixel,
                             n, s);
      break;
    }
    case kStepCopy:
      launch_copy(pi.base, pi.tile_stride, po.base + st.copy_dst_off, po.tile_stride, st.copy_bytes, n, s);
      break;
  }
  return TOD_OK;
}

// default YOLACT priors of the FRC head (upstream make_priors order: level, row, column, aspect ratio)
int default_priors(int P, std::vector<float>* out) {
  static const int kSizes[5] = {28, 14, 7, 4, 2};
  static const float kScales[5] = {12.0f, 24.0f, 48.0f, 96.0f, 192.0f};
  static const float kAspect[3] = {1.0f, 0.5f, 2.0f};
  int total = 0;
  for (int s : kSizes) total += 3 * s * s;
  if (total != P) return -1;
  out->clear();
  for (int l = 0; l < 5; ++l) {
    const int s = kSizes[l];
    for (int j = 0; j < s; ++j)
      for (int i = 0; i < s; ++i) {
        const float x = (float(i) + 0.5f) / float(s), yv = (float(j) + 0.5f) / float(s);
        for (int a = 0; a < 3; ++a) {
          const float ar = std::sqrt(kAspect[a]);
          out->push_back(x);
          out->push_back(yv);
          out->push_back(kScales[l] * ar / 224.0f);
          out->push_back(kScales[l] / ar / 224.0f);
        }
      }
  }
  return 0;
}

template <class T>
int dev_alloc(tod_yolact* y, T** p, size_t count) {
  void* v = nullptr;
  TOD_CUDA(cudaMalloc(&v, count * sizeof(T) ? count * sizeof(T) : 256));
  y->det_allocs.push_back(v);
  *p = static_cast<T*>(v);
  return TOD_OK;
}

int setup_detection(tod_yolact* y) {
  const Graph& G = y->graph;
  if (G.outputs.size() < 5) return TOD_OK;
  // outputs: spatial tensors [1,h,w,c] (seg == outputs[4] as in yolact.rs:91, proto == the other one)
  // and head tensors [1,P,c] (box: c == 4, coef: c == proto channels, cls: the remaining one)
  std::vector<int> spatial, head;
  for (size_t i = 0; i < G.outputs.size(); ++i) {
    const GTensor& X = G.tensors[G.outputs[i]];
    if (X.type != kU8) return TOD_OK;
    if (X.dims[1] > 1) spatial.push_back(int(i));
    else head.push_back(int(i));
  }
  if (spatial.size() != 2 || head.size() != 3) return TOD_OK;
  const int seg_i = 4;
  if (std::find(spatial.begin(), spatial.end(), seg_i) == spatial.end()) return TOD_OK;
  const int proto_i = spatial[0] == seg_i ? spatial[1] : spatial[0];
  const GTensor& PR = G.tensors[G.outputs[proto_i]];
  int box_i = -1, coef_i = -1, cls_i = -1;
  for (int h : head) {
    const GTensor& X = G.tensors[G.outputs[h]];
    if (X.dims[3] == 4 && box_i < 0) box_i = h;
    else if (X.dims[3] == PR.dims[3] && coef_i < 0) coef_i = h;
    else cls_i = h;
  }
  if (box_i < 0 || coef_i < 0 || cls_i < 0) return TOD_OK;
  const GTensor& BX = G.tensors[G.outputs[box_i]];
  const GTensor& CL = G.tensors[G.outputs[cls_i]];
  const GTensor& CF = G.tensors[G.outputs[coef_i]];
  if (BX.dims[2] != CL.dims[2] || BX.dims[2] != CF.dims[2]) return TOD_OK;
  DetectCfg c{};
  c.P = BX.dims[2];
  c.C = CL.dims[3];
  c.K = CF.dims[3];
  c.ph = PR.dims[1];
  c.pw = PR.dims[2];
  c.conf_thresh = y->opt.conf_thresh;
  c.nms_thresh = y->opt.nms_thresh;
  c.top_k = y->opt.top_k;
  c.max_dets = y->opt.max_dets;
  c.box_zp = BX.zp();
  c.coef_zp = CF.zp();
  c.proto_zp = PR.zp();
  c.mask_scale = PR.scale() * CF.scale();
  if (c.C < 2 || c.top_k < 1 || c.max_dets < 1) return TOD_OK;
  TOD_TRY(detect_setup_kernels(c));
  y->o_box = box_i; y->o_cls = cls_i; y->o_coef = coef_i; y->o_proto = proto_i;
  y->dcfg = c;
  const size_t mt = size_t(y->opt.max_tiles), fg = size_t(c.C - 1);
  DetectBuffers& b = y->dbuf;
  float *exp_diff, *box_deq, *box_exp;
  TOD_TRY(dev_alloc(y, &y->d_priors, size_t(c.P) * 4));
  TOD_TRY(dev_alloc(y, &exp_diff, 256));
  TOD_TRY(dev_alloc(y, &box_deq, 256));
  TOD_TRY(dev_alloc(y, &box_exp, 256));
  TOD_TRY(dev_alloc(y, &b.boxes, mt * c.P * 4));
  TOD_TRY(dev_alloc(y, &b.cand, mt * fg * c.P));
  TOD_TRY(dev_alloc(y, &b.cand_count, mt * fg));
  TOD_TRY(dev_alloc(y, &b.surv, mt * fg * c.top_k));
  TOD_TRY(dev_alloc(y, &b.surv_count, mt));
  TOD_TRY(dev_alloc(y, &b.det_count, mt));
  TOD_TRY(dev_alloc(y, &b.det_box, mt * c.max_dets * 4));
  TOD_TRY(dev_alloc(y, &b.det_score, mt * c.max_dets));
  TOD_TRY(dev_alloc(y, &b.det_class, mt * c.max_dets));
  TOD_TRY(dev_alloc(y, &b.det_prior, mt * c.max_dets));
  TOD_TRY(dev_alloc(y, &b.masks, mt * c.max_dets * c.ph * c.pw));
  TOD_TRY(dev_alloc(y, &b.masks_bin, mt * c.max_dets * c.ph * c.pw));
  TOD_TRY(dev_alloc(y, &b.masks_bits, mt * c.max_dets * size_t((c.ph * c.pw + 31) / 32)));
  // exp() of a quantised logit is a function of the u8 code only: 256-entry tables built with the host libm
  float h_exp[256], h_deq[256], h_bexp[256];
  for (int d = 0; d < 256; ++d) h_exp[d] = std::exp(CL.scale() * float(d - 255));
  for (int q = 0; q < 256; ++q) {
    h_deq[q] = BX.scale() * float(q - c.box_zp);
    h_bexp[q] = std::exp(h_deq[q] * 0.2f);
  }
  TOD_CUDA(cudaMemcpy(exp_diff, h_exp, sizeof(h_exp), cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(box_deq, h_deq, sizeof(h_deq), cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(box_exp, h_bexp, sizeof(h_bexp), cudaMemcpyHostToDevice));
  b.priors = y->d_priors;
  b.exp_diff = exp_diff;
  b.box_deq = box_deq;
  b.box_exp = box_exp;
  y->det_ready = true;
  std::vector<float> pri;
  if (default_priors(c.P, &pri) == 0) {
    TOD_CUDA(cudaMemcpy(y->d_priors, pri.data(), pri.size() * 4, cudaMemcpyHostToDevice));
    y->have_priors = true;
  }
  return TOD_OK;
}

int enqueue_post(tod_yolact* y, int n, bool dets, int masks, cudaStream_t s) {
  if (y->seg_out >= 0) {
    const GTensor& S = y->T(y->graph.outputs[y->seg_out]);
    const Place& ps = y->place[y->graph.outputs[y->seg_out]];
    SegPost p{S.dims[1], S.dims[2], S.dims[3], S.scale(), S.zp(), y->opt.id_mode, y->tile_w() / S.dims[2]};
    launch_seg_postprocess(ps.base, ps.tile_stride, n, p, y->d_tile_classes, y->d_cell_classes, y->d_diverges, s);
  }
  if (dets) {
    const Place& pc = y->place[y->graph.outputs[y->o_cls]];
    const Place& pb = y->place[y->graph.outputs[y->o_box]];
    const Place& pf = y->place[y->graph.outputs[y->o_coef]];
    const Place& pp = y->place[y->graph.outputs[y->o_proto]];
    DetectBuffers db = y->dbuf;
    if (masks < 2) db.masks = nullptr;  // float masks only when somebody will read them
    TOD_TRY(launch_detect(y->dcfg, db, pc.base, pc.tile_stride, pb.base, pb.tile_stride, pf.base, pf.tile_stride, pp.base,
                          pp.tile_stride, n, masks != 0, s));
  }
  return TOD_OK;
}

// the per-batch pipeline after the input tiles are in place: graph steps + post-processing
int enqueue_all(tod_yolact* y, int n, bool dets, int masks, cudaStream_t s) {
  for (const Step& st : y->steps) TOD_TRY(run_step(y, st, n, s));
  TOD_TRY(enqueue_post(y, n, dets, masks, s));
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

// Same work as enqueue_all, spread over the handle's lane streams by data dependency.  Only used while capturing:
// the resulting CUDA graph has one node per kernel and an edge per dependency, so independent branches overlap.
int enqueue_all_parallel(tod_yolact* y, int n, bool dets, int masks, cudaStream_t origin) {
  const int L = tod_yolact::kLanes;
  std::vector<int> lane_of(y->steps.size(), 0), tail(L, -1);
  std::vector<char> forked(L, 0);
  std::vector<std::vector<char>> ancestors(y->steps.size());
  if (!y->trace_events.empty()) TOD_CUDA(cudaEventRecord(y->trace_events[y->steps.size()], origin));
  TOD_CUDA(cudaEventRecord(y->fork_event, origin));
  for (size_t i = 0; i < y->steps.size(); ++i) {
    const Step& st = y->steps[i];
    // Stream order becomes a graph edge, so a step only joins a lane whose tail it depends on anyway: first the lane
    // of a direct producer that is still a tail (a kernel -> kernel chain, PDL-eligible), else a lane whose tail is one
    // of its ancestors, else an unused lane; the captured graph then holds the true dependencies only (with six
    // round-robin lanes the small pyramid levels' heads queued behind unrelated layers and held detection back ~250 us).
    std::vector<char>& anc = ancestors[i];
    anc.assign(y->steps.size(), 0);
    for (int d : st.deps) {
      anc[d] = 1;
      for (size_t k = 0; k < y->steps.size(); ++k) anc[k] |= ancestors[d][k];
    }
    int lane = -1;
    for (auto it = st.deps.rbegin(); it != st.deps.rend() && lane < 0; ++it)
      if (tail[lane_of[*it]] == *it) lane = lane_of[*it];
    for (int l = 0; l < L && lane < 0; ++l)
      if (tail[l] >= 0 && anc[tail[l]]) lane = l;
    for (int l = 0; l < L && lane < 0; ++l)
      if (tail[l] < 0) lane = l;
    if (lane < 0) {
      lane = 0;
      for (int l = 1; l < L; ++l)
        if (tail[l] < tail[lane]) lane = l;  // out of lanes: least recently used
    }
    cudaStream_t ls = y->lanes[lane];
    if (!forked[lane]) {
      TOD_CUDA(cudaStreamWaitEvent(ls, y->fork_event, 0));
      forked[lane] = 1;
    }
    bool waited = false;
    for (int d : st.deps)
      if (lane_of[d] != lane) {
        TOD_CUDA(cudaStreamWaitEvent(ls, y->step_events[d], 0));
        waited = true;
      }
    // programmatic dependent launch only along an unbroken kernel -> kernel chain of this lane: the previous node of
    // the stream must be the kernel this step waits for (event records in between are fine, event waits are not)
    pdl_next() = y->opt.use_pdl && !waited && tail[lane] >= 0 && !st.deps.empty();  // only launch_k kernels honour it, and all of those wait
    const int rc_step = run_step(y, st, n, ls);
    pdl_next() = false;
    TOD_TRY(rc_step);
    TOD_CUDA(cudaEventRecord(y->step_events[i], ls));
    if (!y->trace_events.empty()) {
      TOD_CUDA(cudaEventRecord(y->trace_events[i], ls));
      y->trace_lanes[i] = lane;
    }
    lane_of[i] = lane;
    tail[lane] = int(i);
  }
  // post-processing rides the branches: the literal segmentation pass follows the seg head on its lane; box decode /
  // Fast-NMS / top-k follow the class and box heads on theirs and so overlap the protonet branch; only mask assembly
  // (prototypes x coefficients of the selected detections) waits for everything.
  auto after_outputs = [&](std::initializer_list<int> outs, int* lane_out) -> int {
    int last = -1;
    for (int o : outs)
      for (int sidx : y->out_steps[o]) last = std::max(last, sidx);
    const int lane = last >= 0 ? lane_of[last] : 0;
    cudaStream_t ls = y->lanes[lane];
    if (!forked[lane]) {
      TOD_CUDA(cudaStreamWaitEvent(ls, y->fork_event, 0));
      forked[lane] = 1;
    }
    for (int o : outs)
      for (int sidx : y->out_steps[o])
        if (lane_of[sidx] != lane || sidx != tail[lane]) TOD_CUDA(cudaStreamWaitEvent(ls, y->step_events[sidx], 0));
    *lane_out = lane;
    return TOD_OK;
  };
  bool seg_done = false, det_done = false;
  if (y->seg_out >= 0) {
    int lane = 0;
    TOD_TRY(after_outputs({y->seg_out}, &lane));
    const GTensor& S = y->T(y->graph.outputs[y->seg_out]);
    const Place& ps = y->place[y->graph.outputs[y->seg_out]];
    SegPost p{S.dims[1], S.dims[2], S.dims[3], S.scale(), S.zp(), y->opt.id_mode, y->tile_w() / S.dims[2]};
    launch_seg_postprocess(ps.base, ps.tile_stride, n, p, y->d_tile_classes, y->d_cell_classes, y->d_diverges, y->lanes[lane]);
    // an external event node: the host-facing call starts copying the class maps back as soon as they exist
    if (y->trace_events.empty()) TOD_CUDA(cudaEventRecordWithFlags(y->seg_ready, y->lanes[lane], cudaEventRecordExternal));
    else TOD_CUDA(cudaEventRecord(y->seg_ready, y->lanes[lane]));  // tod_yolact_trace_steps runs outside a capture
    TOD_CUDA(cudaEventRecord(y->post_events[0], y->lanes[lane]));
    if (!y->trace_events.empty()) TOD_CUDA(cudaEventRecord(y->trace_events[y->steps.size() + 1], y->lanes[lane]));
    seg_done = true;
  }
  if (dets) {
    int lane = 0;
    TOD_TRY(after_outputs({y->o_cls, y->o_box}, &lane));
    const Place& pc = y->place[y->graph.outputs[y->o_cls]];
    const Place& pb = y->place[y->graph.outputs[y->o_box]];
    TOD_TRY(launch_detect_boxes(y->dcfg, y->dbuf, pc.base, pc.tile_stride, pb.base, pb.tile_stride, n, y->lanes[lane], y->nms_stream, y->nms_fork, y->nms_join));
    TOD_CUDA(cudaEventRecord(y->post_events[1], y->lanes[lane]));
    if (!y->trace_events.empty()) TOD_CUDA(cudaEventRecord(y->trace_events[y->steps.size() + 2], y->lanes[lane]));
    det_done = true;
  }
  for (int l = 0; l < L; ++l)
    if (forked[l] && tail[l] >= 0) TOD_CUDA(cudaStreamWaitEvent(origin, y->step_events[tail[l]], 0));
  if (seg_done) TOD_CUDA(cudaStreamWaitEvent(origin, y->post_events[0], 0));
  if (det_done) TOD_CUDA(cudaStreamWaitEvent(origin, y->post_events[1], 0));
  if (dets && masks) {
    const Place& pf = y->place[y->graph.outputs[y->o_coef]];
    const Place& pp = y->place[y->graph.outputs[y->o_proto]];
    DetectBuffers db = y->dbuf;
    if (masks < 2) db.masks = nullptr;
    TOD_TRY(launch_detect_masks(y->dcfg, db, pf.base, pf.tile_stride, pp.base, pp.tile_stride, n, origin));
  }
  if (!y->trace_events.empty()) TOD_CUDA(cudaEventRecord(y->trace_events[y->steps.size() + 3], origin));
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

int run_pipeline(tod_yolact* y, int n, bool dets, int masks, cudaStream_t s) {
  if (dets && !y->det_ready) return fail(TOD_ERR_UNSUPPORTED, "this model's outputs do not form a YOLACT detection head");
  if (dets && !y->have_priors) return fail(TOD_ERR_INVALID_ARG, "no priors for a %d-prior head: call tod_yolact_set_priors first", y->dcfg.P);
  y->last_tiles = n;
  if (!y->opt.use_cuda_graph) return enqueue_all(y, n, dets, masks, s);
  const int key = (n << 3) | (dets ? 4 : 0) | (masks & 3);
  auto it = y->graphs.find(key);
  if (it == y->graphs.end()) {
    cudaGraph_t g = nullptr;
    TOD_CUDA(cudaStreamBeginCapture(y->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_all_parallel(y, n, dets, masks, y->stream);
    const cudaError_t ce = cudaStreamEndCapture(y->stream, &g);
    if (rc < 0) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (ce != cudaSuccess) return fail(TOD_ERR_CUDA, "CUDA graph capture failed: %s", cudaGetErrorString(ce));
    cudaGraphExec_t ge = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) return fail(TOD_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
    it = y->graphs.emplace(key, ge).first;
  }
  TOD_CUDA(cudaGraphLaunch(it->second, s));
  return TOD_OK;
}

int fetch_strided(void* dst, const Place& p, int n, cudaStream_t s) {
  if (p.c_store != p.c && p.c > 0) {  // channel-padded storage: copy the real channels of every pixel
    const size_t pixels = size_t(p.bytes / p.c_store);
    for (int t = 0; t < n; ++t)
      TOD_CUDA(cudaMemcpy2DAsync(static_cast<uint8_t*>(dst) + size_t(t) * pixels * p.c, size_t(p.c), p.base + size_t(t) * p.tile_stride, size_t(p.c_store),
                                 size_t(p.c), pixels, cudaMemcpyDeviceToHost, s));
    return TOD_OK;
  }
  TOD_CUDA(cudaMemcpy2DAsync(dst, size_t(p.bytes), p.base, size_t(p.tile_stride), size_t(p.bytes), size_t(n), cudaMemcpyDeviceToHost, s));
  return TOD_OK;
}

int upload_axis(AxisTables* t) {
  cudaFree(t->d_left); cudaFree(t->d_count); cudaFree(t->d_weight);
  t->d_left = nullptr; t->d_count = nullptr; t->d_weight = nullptr;
  TOD_CUDA(cudaMalloc(&t->d_left, t->left.size() * 4));
  TOD_CUDA(cudaMalloc(&t->d_count, t->count.size() * 4));
  TOD_CUDA(cudaMalloc(&t->d_weight, t->weight.size() * 4));
  TOD_CUDA(cudaMemcpy(t->d_left, t->left.data(), t->left.size() * 4, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(t->d_count, t->count.data(), t->count.size() * 4, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(t->d_weight, t->weight.data(), t->weight.size() * 4, cudaMemcpyHostToDevice));
  return TOD_OK;
}

ResampleAxis axis_of(const AxisTables& t) { return ResampleAxis{t.d_left, t.d_count, t.d_weight, t.in_size, t.out_size}; }

int ensure_classify_buffers(tod_yolact* y, int W, int H) {
  if (y->cw == W && y->chh == H) return TOD_OK;
  if (W < 2 || H < 2 || W > 8192 || H > 8192) return fail(TOD_ERR_INVALID_ARG, "classify: unsupported frame size %dx%d", W, H);
  y->cw = y->chh = 0;   // nothing below is valid for the old size any more; set again only when everything succeeded
  const int tw = y->tile_w(), th = y->tile_h();
  TOD_TRY(build_axis(H, th, &y->pre_v));       // 640x480 -> 448x224 (yolact.rs:208): vertical pass first
  TOD_TRY(build_axis(W, 2 * tw, &y->pre_h));
  TOD_TRY(build_axis(th, H, &y->post_v));      // 448x224 -> 640x480 (yolact.rs:231)
  TOD_TRY(build_axis(2 * tw, W, &y->post_h));
  TOD_TRY(upload_axis(&y->pre_v));
  TOD_TRY(upload_axis(&y->pre_h));
  TOD_TRY(upload_axis(&y->post_v));
  TOD_TRY(upload_axis(&y->post_h));
  const size_t frames = size_t(y->opt.max_tiles) / 2;
  cudaFree(y->d_tmp); cudaFree(y->d_frames); cudaFree(y->d_tiles_rgb);
  y->d_tmp = nullptr; y->d_frames = nullptr; y->d_tiles_rgb = nullptr;
  const size_t tmp_px = std::max(size_t(th) * W, size_t(H) * 2 * tw);
  TOD_CUDA(cudaMalloc(&y->d_tmp, frames * tmp_px * 3 * sizeof(float)));
  TOD_CUDA(cudaMalloc(&y->d_frames, frames * size_t(W) * H * 4));
  TOD_CUDA(cudaMalloc(&y->d_tiles_rgb, frames * 2 * size_t(tw) * th * 3));
  y->cw = W;
  y->chh = H;
  return TOD_OK;
}

int classify_device(tod_yolact* y, uint32_t* d_frames, int n, int W, int H, uint16_t* d_target, cudaStream_t s) {
  const int tw = y->tile_w(), th = y->tile_h();
  if (y->seg_out < 0) return fail(TOD_ERR_UNSUPPORTED, "classify needs output #4 ([1,h,w,c] segmentation logits, yolact.rs:91)");
  launch_classify_pre(d_frames, n, W, H, axis_of(y->pre_v), axis_of(y->pre_h), y->d_tmp, y->d_tiles_rgb, tw, th, s);
  const Place& pin = y->place[y->graph.inputs[0]];
  TOD_CUDA(cudaMemcpy2DAsync(pin.base, size_t(pin.tile_stride), y->d_tiles_rgb, size_t(pin.bytes), size_t(pin.bytes), size_t(2 * n),
                             cudaMemcpyDeviceToDevice, s));
  TOD_TRY(run_pipeline(y, 2 * n, false, 0, s));
  launch_classify_post(y->d_tile_classes, n, W, H, axis_of(y->post_v), axis_of(y->post_h), y->d_tmp, d_frames, d_target, tw, th, s);
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

// Host-facing calls block like the reference's (`invoke()`, `future.wait`).  The wait sleeps on a blocking-sync event
// instead of spinning in cudaStreamSynchronize: a frame loop keeps several handles / GPUs busy from a few host threads,
// and a spinning waiter per handle starves them once threads outnumber cores (8 ranks x 3 handles on 16 cores).
int wait_stream(tod_yolact* y, cudaStream_t s, cudaEvent_t ev) {
  if (y->spin_sync) {
    TOD_CUDA(cudaStreamSynchronize(s));
    return TOD_OK;
  }
  TOD_CUDA(cudaEventRecord(ev, s));
  // poll with sched_yield for the first few milliseconds (a step is ~1 ms: no interrupt wake-up latency, and a core that
  // other runnable threads want is handed over at once), then sleep on the blocking-sync event
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t q = cudaEventQuery(ev);
    if (q == cudaSuccess) return TOD_OK;
    if (q != cudaErrorNotReady) return fail(TOD_ERR_CUDA, "cudaEventQuery failed: %s", cudaGetErrorString(q));
    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(4)) break;
    std::this_thread::yield();
  }
  TOD_CUDA(cudaEventSynchronize(ev));
  return TOD_OK;
}

// A device-resident call may run on a caller's stream; tod_yolact_fetch_* copy on the handle's own stream, which is made
// to follow the caller's work here (one handle still serves one stream at a time: its scratch buffers are shared).
int order_after_caller(tod_yolact* y, cudaStream_t s) {
  if (s == y->stream) return TOD_OK;
  TOD_CUDA(cudaEventRecord(y->caller_done, s));
  TOD_CUDA(cudaStreamWaitEvent(y->stream, y->caller_done, 0));
  return TOD_OK;
}

int enqueue_diverged(tod_yolact* y, int tiles, cudaStream_t s) {
  if (y->opt.id_mode != 0 || y->seg_out < 0) return TOD_OK;
  TOD_CUDA(cudaMemcpyAsync(y->h_diverges, y->d_diverges, sizeof(int) * tiles, cudaMemcpyDeviceToHost, s));
  return TOD_OK;
}

int read_diverged(const tod_yolact* y, int tiles) {
  if (y->opt.id_mode != 0 || y->seg_out < 0) return TOD_OK;
  for (int i = 0; i < tiles; ++i)
    if (y->h_diverges[i]) return TOD_WARN_REFERENCE_DIVERGES;
  return TOD_OK;
}

int check_diverged(tod_yolact* y, int tiles, cudaStream_t s) {
  if (y->opt.id_mode != 0 || y->seg_out < 0) return TOD_OK;
  TOD_TRY(enqueue_diverged(y, tiles, s));
  TOD_TRY(wait_stream(y, s, y->done_ev));
  return read_diverged(y, tiles);
}

}  // namespace

extern "C" {

void tod_yolact_default_options(tod_yolact_options* o) {
  if (!o) return;
  o->max_tiles = 2;
  o->id_mode = 0;
  o->conf_thresh = 0.05f;
  o->nms_thresh = 0.5f;
  o->top_k = 200;
  o->max_dets = 100;
  o->use_cuda_graph = 1;
  o->conv_impl = 0;
  o->fusion = 1;
  o->use_pdl = 1;
  o->batches_in_flight = 1;
}

int tod_model_inspect(const char* tflite_path, int32_t* num_ops, int32_t* num_tensors, int64_t* macs) {
  Graph g;
  TOD_TRY(read_tflite(tflite_path, &g));
  if (num_ops) *num_ops = int32_t(g.ops.size());
  if (num_tensors) *num_tensors = int32_t(g.tensors.size());
  if (macs) {
    int64_t m = 0;
    for (const GOp& op : g.ops) {
      if (op.code != kConv2D && op.code != kDepthwise) continue;
      if (op.inputs.size() < 2 || op.outputs.empty()) continue;
      const GTensor& W = g.tensors[op.inputs[1]];
      const GTensor& O = g.tensors[op.outputs[0]];
      const int64_t per = op.code == kConv2D ? int64_t(W.dims[1]) * W.dims[2] * W.dims[3] : int64_t(W.dims[1]) * W.dims[2];
      m += O.elems() * per;
    }
    *macs = m;
  }
  return TOD_OK;
}

void tod_yolact_destroy(tod_yolact* y) {
  if (!y) return;
  cudaSetDevice(y->device);
  if (y->stream) cudaStreamSynchronize(y->stream);
  for (auto& kv : y->graphs) cudaGraphExecDestroy(kv.second);
  for (cudaStream_t l : y->lanes)
    if (l) cudaStreamDestroy(l);
  for (cudaEvent_t e : y->step_events)
    if (e) cudaEventDestroy(e);
  if (y->fork_event) cudaEventDestroy(y->fork_event);
  for (cudaEvent_t e : y->post_events)
    if (e) cudaEventDestroy(e);
  for (Step& s : y->steps)
    if (s.tc) conv_tc_destroy(s.tc);
  for (void* p : y->det_allocs) cudaFree(p);
  for (AxisTables* t : {&y->pre_v, &y->pre_h, &y->post_v, &y->post_h}) {
    cudaFree(t->d_left); cudaFree(t->d_count); cudaFree(t->d_weight);
  }
  cudaFree(y->d_tmp); cudaFree(y->d_frames); cudaFree(y->d_tiles_rgb);
  cudaFree(y->d_tile_classes); cudaFree(y->d_cell_classes); cudaFree(y->d_diverges); cudaFree(y->d_tile_bits);
  if (y->caller_done) cudaEventDestroy(y->caller_done);
  if (y->done_ev) cudaEventDestroy(y->done_ev);
  if (y->done_ev2) cudaEventDestroy(y->done_ev2);
  if (y->h_diverges) cudaFreeHost(y->h_diverges);
  cudaFree(y->d_const); cudaFree(y->d_act);
  if (y->copy_stream) cudaStreamDestroy(y->copy_stream);
  if (y->seg_ready) cudaEventDestroy(y->seg_ready);
  if (y->nms_fork) cudaEventDestroy(y->nms_fork);
  if (y->nms_join) cudaEventDestroy(y->nms_join);
  if (y->nms_stream) cudaStreamDestroy(y->nms_stream);
  if (y->stream) cudaStreamDestroy(y->stream);
  delete y;
}

int tod_yolact_create(const char* tflite_path, int device, const tod_yolact_options* opts, tod_yolact** out) {
  if (!tflite_path || !out) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_create: null argument");
  *out = nullptr;
  tod_yolact_options o;
  tod_yolact_default_options(&o);
  if (opts) o = *opts;
  if (o.max_tiles < 1 || o.max_tiles > 16384) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_create: max_tiles must be in [1,16384]");
  if (o.id_mode != 0 && o.id_mode != 1) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_create: id_mode must be 0 or 1");
  std::unique_ptr<tod_yolact> y(new tod_yolact());
  y->device = device;
  y->opt = o;
  TOD_TRY(read_tflite(tflite_path, &y->graph));  // before touching the GPU: a bad model fails the same way everywhere
  TOD_TRY(select_device(device));
  tod_yolact* raw = y.release();
  auto bail = [&](int rc) {
    tod_yolact_destroy(raw);
    return rc;
  };
  cudaError_t ce = cudaStreamCreateWithFlags(&raw->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&raw->copy_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&raw->seg_ready, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&raw->nms_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&raw->nms_fork, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&raw->nms_join, cudaEventDisableTiming);
  if (ce != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)));
  ConstArena arena;
  int rc = plan(raw, &arena);
  if (rc < 0) return bail(rc);
  rc = upload_consts_and_bind(raw, arena);
  if (rc < 0) return bail(rc);
  // literal post-processing needs output #4 == [1,gh,gw,c] (yolact.rs:91,108,118)
  const Graph& G = raw->graph;
  if (G.outputs.size() > 4) {
    const GTensor& S = G.tensors[G.outputs[4]];
    if (S.type == kU8 && S.dims[1] > 1 && S.dims[1] * S.dims[2] <= 1024 && S.dims[3] >= 4 && raw->tile_w() % S.dims[2] == 0 &&
        raw->tile_h() % S.dims[1] == 0 && raw->tile_w() / S.dims[2] == raw->tile_h() / S.dims[1])
      raw->seg_out = 4;
  }
  const size_t mt = size_t(o.max_tiles);
  raw->diag_skip = std::getenv("TOD_DIAG_SKIP") ? std::atoi(std::getenv("TOD_DIAG_SKIP")) : 0;
  raw->spin_sync = std::getenv("TOD_SPIN_SYNC") && std::atoi(std::getenv("TOD_SPIN_SYNC")) != 0;
  if ((ce = cudaEventCreateWithFlags(&raw->caller_done, cudaEventDisableTiming)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&raw->done_ev, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&raw->done_ev2, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess)
    return bail(fail(TOD_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
  if ((ce = cudaMalloc(&raw->d_cell_classes, mt * 1024 * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&raw->d_tile_classes, mt * raw->tile_w() * raw->tile_h() * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&raw->d_diverges, mt * sizeof(int))) != cudaSuccess ||
      (ce = cudaMallocHost(&raw->h_diverges, mt * sizeof(int))) != cudaSuccess)
    return bail(fail(TOD_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(ce)));
  cudaMemset(raw->d_diverges, 0, mt * sizeof(int));
  rc = setup_detection(raw);
  if (rc < 0) return bail(rc);
  raw->launches_per_call = int(raw->steps.size());
  for (cudaStream_t& l : raw->lanes)
    if ((ce = cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking)) != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce)));
  raw->step_events.assign(raw->steps.size(), nullptr);
  for (cudaEvent_t& e : raw->step_events)
    if ((ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming)) != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
  if ((ce = cudaEventCreateWithFlags(&raw->fork_event, cudaEventDisableTiming)) != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
  for (cudaEvent_t& e : raw->post_events)
    if ((ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming)) != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce)));
  *out = raw;
  return TOD_OK;
}

int tod_yolact_set_priors(tod_yolact* y, const float* priors, int n) {
  if (!y || !priors) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_set_priors: null argument");
  if (!y->det_ready) return fail(TOD_ERR_UNSUPPORTED, "this model's outputs do not form a YOLACT detection head");
  if (n != y->dcfg.P) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_set_priors: the head has %d priors, got %d", y->dcfg.P, n);
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_CUDA(cudaMemcpy(y->d_priors, priors, size_t(n) * 16, cudaMemcpyHostToDevice));
  y->have_priors = true;
  return TOD_OK;
}

int tod_yolact_num_outputs(const tod_yolact* y) { return y ? int(y->graph.outputs.size()) : 0; }
int tod_yolact_num_tensors(const tod_yolact* y) { return y ? int(y->graph.tensors.size()) : 0; }
int tod_yolact_num_ops(const tod_yolact* y) { return y ? int(y->graph.ops.size()) : 0; }

int tod_yolact_tensor_info(const tod_yolact* y, int tensor, int32_t shape4[4], int32_t* type, float* scale, int32_t* zero_point,
                           int32_t* elems) {
  if (!y || tensor < 0 || tensor >= int(y->graph.tensors.size())) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_tensor_info: bad tensor index");
  const GTensor& X = y->graph.tensors[tensor];
  if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = X.dims[i];
  if (type) *type = X.type;
  if (scale) *scale = X.scale();
  if (zero_point) *zero_point = X.zp();
  if (elems) *elems = int32_t(X.elems());
  return TOD_OK;
}

int tod_yolact_input_info(const tod_yolact* y, int32_t shape4[4]) {
  if (!y || !shape4) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_input_info: null argument");
  return tod_yolact_tensor_info(y, y->graph.inputs[0], shape4, nullptr, nullptr, nullptr, nullptr);
}

int tod_yolact_output_info(const tod_yolact* y, int index, int32_t shape4[4], float* scale, int32_t* zero_point, int32_t* elems) {
  if (!y || index < 0 || index >= int(y->graph.outputs.size())) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_output_info: bad output index");
  return tod_yolact_tensor_info(y, y->graph.outputs[index], shape4, nullptr, scale, zero_point, elems);
}

int tod_yolact_infer_tiles_device(tod_yolact* y, const uint8_t* d_rgb_tiles, int n, void* stream) {
  if (!y || !d_rgb_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_infer_tiles_device: null argument");
  if (n < 1 || n > y->opt.max_tiles) return fail(TOD_ERR_CAPACITY, "tod_yolact_infer_tiles_device: n=%d outside [1,%d]", n, y->opt.max_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : y->stream;
  const Place& pin = y->place[y->graph.inputs[0]];
  TOD_CUDA(cudaMemcpy2DAsync(pin.base, size_t(pin.tile_stride), d_rgb_tiles, size_t(pin.bytes), size_t(pin.bytes), size_t(n),
                             cudaMemcpyDeviceToDevice, s));
  const bool diag_nodets = (y->diag_skip & 8) != 0;
  const bool dets = y->det_ready && y->have_priors && !diag_nodets;
  y->last_mask_mode = dets ? 2 : 0;
  TOD_TRY(run_pipeline(y, n, dets, dets ? 2 : 0, s));
  return order_after_caller(y, s);
}

int tod_yolact_fetch_output(tod_yolact* y, int index, int n, uint8_t* out) {
  if (!y || !out || index < 0 || index >= int(y->graph.outputs.size())) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_output: bad argument");
  if (n < 1 || n > y->last_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_output: n=%d but the last call ran %d tiles", n, y->last_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_TRY(fetch_strided(out, y->place[y->graph.outputs[index]], n, y->stream));
  TOD_CUDA(cudaStreamSynchronize(y->stream));
  return TOD_OK;
}

int tod_yolact_fetch_output_f32(tod_yolact* y, int index, int n, float* out) {
  if (!y || !out || index < 0 || index >= int(y->graph.outputs.size())) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_output_f32: bad argument");
  if (n < 1 || n > y->last_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_output_f32: n=%d but the last call ran %d tiles", n, y->last_tiles);
  const GTensor& T = y->T(y->graph.outputs[index]);
  if (T.type != kU8) return fail(TOD_ERR_UNSUPPORTED, "tod_yolact_fetch_output_f32: output %d is not uint8", index);
  const Place& p = y->place[y->graph.outputs[index]];
  const int c = p.c_store > 0 ? p.c : 1, c_store = p.c_store > 0 ? p.c_store : 1;
  const int64_t elems = p.c_store > 0 ? p.bytes / p.c_store * p.c : p.bytes;
  TOD_CUDA(cudaSetDevice(y->device));
  float* d_f = nullptr;
  TOD_CUDA(cudaMalloc(&d_f, size_t(n) * elems * 4));
  int rc = launch_dequant_u8(p.base, p.tile_stride, c, c_store, elems, n, T.scale(), T.zp(), d_f, y->stream);
  if (rc == TOD_OK && cudaMemcpyAsync(out, d_f, size_t(n) * elems * 4, cudaMemcpyDeviceToHost, y->stream) != cudaSuccess) rc = fail(TOD_ERR_CUDA, "tod_yolact_fetch_output_f32: copy failed");
  if (cudaStreamSynchronize(y->stream) != cudaSuccess && rc == TOD_OK) rc = fail(TOD_ERR_CUDA, "tod_yolact_fetch_output_f32: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_f);
  return rc;
}

int tod_yolact_fetch_tensor(tod_yolact* y, int tensor, int n, void* out, size_t out_bytes) {
  if (!y || !out || tensor < 0 || tensor >= int(y->graph.tensors.size())) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_tensor: bad argument");
  if (n < 1 || n > y->last_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_tensor: n=%d but the last call ran %d tiles", n, y->last_tiles);
  const Place& p = y->place[tensor];
  if (!p.base) return fail(TOD_ERR_INVALID_ARG, "tensor %d is a constant or is not materialised", tensor);
  const int64_t real_bytes = p.c_store > 0 ? p.bytes / p.c_store * p.c : p.bytes;
  if (out_bytes < size_t(real_bytes) * n) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_tensor: buffer too small (%zu < %lld)", out_bytes, (long long)(real_bytes * n));
  if (!y->fused_away.empty() && y->fused_away[tensor])
    return fail(TOD_ERR_INVALID_ARG, "tensor %d is fused into the producing convolution's epilogue; create the handle with fusion = 0 to fetch it", tensor);
  // a PAD folded into its convolution is never written
  for (size_t i = 0; i < y->graph.ops.size(); ++i)
    if (y->graph.ops[i].code == kPad && y->graph.ops[i].outputs[0] == tensor) {
      bool present = false;
      for (const Step& s : y->steps) present |= (s.kind == kStepPad && s.op == int(i));
      if (!present) return fail(TOD_ERR_INVALID_ARG, "tensor %d (a PAD output) is fused away; create the handle with fusion = 0 to fetch it", tensor);
    }
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_TRY(fetch_strided(out, p, n, y->stream));
  TOD_CUDA(cudaStreamSynchronize(y->stream));
  return TOD_OK;
}

int tod_yolact_fetch_tile_classes(tod_yolact* y, int n, uint32_t* out) {
  if (!y || !out) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_tile_classes: null argument");
  if (y->seg_out < 0) return fail(TOD_ERR_UNSUPPORTED, "the model has no segmentation output #4");
  if (n < 1 || n > y->last_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_tile_classes: n=%d but the last call ran %d tiles", n, y->last_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_CUDA(cudaMemcpyAsync(out, y->d_tile_classes, size_t(n) * y->tile_w() * y->tile_h() * 4, cudaMemcpyDeviceToHost, y->stream));
  TOD_CUDA(cudaStreamSynchronize(y->stream));
  return check_diverged(y, n, y->stream);
}

int tod_yolact_last_diverged(tod_yolact* y, int* diverged) {
  if (!y || !diverged) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_last_diverged: null argument");
  *diverged = 0;
  if (y->last_tiles < 1) return TOD_OK;
  TOD_CUDA(cudaSetDevice(y->device));
  const int rc = check_diverged(y, y->last_tiles, y->stream);
  if (rc < 0) return rc;
  *diverged = rc == TOD_WARN_REFERENCE_DIVERGES ? 1 : 0;
  return TOD_OK;
}

static int enqueue_fetch_detections(tod_yolact* y, int n, tod_detections* d, cudaStream_t s) {
  if (!y->det_ready) return fail(TOD_ERR_UNSUPPORTED, "this model's outputs do not form a YOLACT detection head");
  if (n < 1 || n > y->last_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_detections: n=%d but the last call ran %d tiles", n, y->last_tiles);
  const DetectCfg& c = y->dcfg;
  if (d->max_dets != c.max_dets) return fail(TOD_ERR_INVALID_ARG, "tod_detections.max_dets=%d, handle was created with %d", d->max_dets, c.max_dets);
  if (d->masks && y->last_mask_mode < 2) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_detections: the last call did not compute float masks");
  if ((d->masks_bin || d->masks_bits) && y->last_mask_mode < 1) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_detections: the last call did not compute masks");
  if (d->masks_tile_bits && y->last_mask_mode < 2) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_detections: tile-resolution masks need the float masks of the last call");
  const size_t nd = size_t(n) * c.max_dets;
  const DetectBuffers& b = y->dbuf;
  if (d->count) TOD_CUDA(cudaMemcpyAsync(d->count, b.det_count, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
  if (d->boxes) TOD_CUDA(cudaMemcpyAsync(d->boxes, b.det_box, nd * 16, cudaMemcpyDeviceToHost, s));
  if (d->scores) TOD_CUDA(cudaMemcpyAsync(d->scores, b.det_score, nd * 4, cudaMemcpyDeviceToHost, s));
  if (d->classes) TOD_CUDA(cudaMemcpyAsync(d->classes, b.det_class, nd * 4, cudaMemcpyDeviceToHost, s));
  if (d->priors) TOD_CUDA(cudaMemcpyAsync(d->priors, b.det_prior, nd * 4, cudaMemcpyDeviceToHost, s));
  if (d->masks) TOD_CUDA(cudaMemcpyAsync(d->masks, b.masks, nd * c.ph * c.pw * 4, cudaMemcpyDeviceToHost, s));
  if (d->masks_bits) TOD_CUDA(cudaMemcpyAsync(d->masks_bits, b.masks_bits, nd * size_t((c.ph * c.pw + 31) / 32) * 4, cudaMemcpyDeviceToHost, s));
  if (d->masks_bin) TOD_CUDA(cudaMemcpyAsync(d->masks_bin, b.masks_bin, nd * c.ph * c.pw, cudaMemcpyDeviceToHost, s));
  if (d->masks_tile_bits) {
    const size_t words = size_t(y->tile_h() * y->tile_w() + 31) / 32;
    if (!y->d_tile_bits) TOD_CUDA(cudaMalloc(&y->d_tile_bits, size_t(y->opt.max_tiles) * c.max_dets * words * 4));
    TOD_TRY(launch_mask_upsample(c, b, n, y->tile_h(), y->tile_w(), y->d_tile_bits, s));
    TOD_CUDA(cudaMemcpyAsync(d->masks_tile_bits, y->d_tile_bits, nd * words * 4, cudaMemcpyDeviceToHost, s));
  }
  return TOD_OK;
}

int tod_yolact_fetch_detections(tod_yolact* y, int n, tod_detections* d) {
  if (!y || !d) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_fetch_detections: null argument");
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_TRY(enqueue_fetch_detections(y, n, d, y->stream));
  return wait_stream(y, y->stream, y->done_ev);
}

int tod_yolact_infer_tiles_cells(tod_yolact* y, const uint8_t* rgb_tiles, int n, uint8_t* const* outputs_u8, uint32_t* tile_classes,
                                 uint32_t* cell_classes, tod_detections* dets) {
  if (!y || !rgb_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_infer_tiles: null argument");
  if (n < 1 || n > y->opt.max_tiles) return fail(TOD_ERR_CAPACITY, "tod_yolact_infer_tiles: n=%d outside [1,%d]", n, y->opt.max_tiles);
  if ((tile_classes || cell_classes) && y->seg_out < 0) return fail(TOD_ERR_UNSUPPORTED, "the model has no segmentation output #4");
  TOD_CUDA(cudaSetDevice(y->device));
  cudaStream_t s = y->stream;
  const Place& pin = y->place[y->graph.inputs[0]];
  // yolact.rs:161-162 copy_from_slice into the input tensor
  TOD_CUDA(cudaMemcpy2DAsync(pin.base, size_t(pin.tile_stride), rgb_tiles, size_t(pin.bytes), size_t(pin.bytes), size_t(n), cudaMemcpyHostToDevice, s));
  const bool want_dets = dets != nullptr;
  const int mask_mode = !want_dets ? 0 : ((dets->masks || dets->masks_tile_bits) ? 2 : ((dets->masks_bin || dets->masks_bits) ? 1 : 0));
  y->last_mask_mode = mask_mode;
  TOD_TRY(run_pipeline(y, n, want_dets, mask_mode, s));
  if (outputs_u8)
    for (size_t k = 0; k < y->graph.outputs.size(); ++k)
      if (outputs_u8[k]) TOD_TRY(fetch_strided(outputs_u8[k], y->place[y->graph.outputs[k]], n, s));
  // the class maps are final as soon as seg_post_kernel has run (the graph records y->seg_ready right behind it):
  // they are read back on a second stream while boxes / NMS / masks are still being computed
  cudaStream_t cs = (y->opt.use_cuda_graph && y->seg_out >= 0) ? y->copy_stream : s;
  const bool side = (tile_classes || cell_classes) && cs != s;
  if (side) TOD_CUDA(cudaStreamWaitEvent(cs, y->seg_ready, 0));
  if (tile_classes) TOD_CUDA(cudaMemcpyAsync(tile_classes, y->d_tile_classes, size_t(n) * y->tile_w() * y->tile_h() * 4, cudaMemcpyDeviceToHost, cs));
  if (cell_classes && y->seg_out >= 0) {
    const GTensor& S = y->T(y->graph.outputs[y->seg_out]);
    TOD_CUDA(cudaMemcpyAsync(cell_classes, y->d_cell_classes, size_t(n) * S.dims[1] * S.dims[2] * 4, cudaMemcpyDeviceToHost, cs));
  }
  if (want_dets) TOD_TRY(enqueue_fetch_detections(y, n, dets, s));
  TOD_TRY(enqueue_diverged(y, n, s));
  // one wait per stream for the whole call
  if (side) TOD_TRY(wait_stream(y, cs, y->done_ev2));
  TOD_TRY(wait_stream(y, s, y->done_ev));
  return read_diverged(y, n);
}

int tod_yolact_infer_tiles(tod_yolact* y, const uint8_t* rgb_tiles, int n, uint8_t* const* outputs_u8, uint32_t* tile_classes,
                           tod_detections* dets) {
  return tod_yolact_infer_tiles_cells(y, rgb_tiles, n, outputs_u8, tile_classes, nullptr, dets);
}

int tod_yolact_classify_batch_device(tod_yolact* y, uint32_t* d_frames, int n, int width, int height, uint16_t* d_target, void* stream) {
  if (!y || !d_frames) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_classify_batch_device: null argument");
  if (n < 1 || 2 * n > y->opt.max_tiles) return fail(TOD_ERR_CAPACITY, "classify: %d frames need %d tiles, handle holds %d", n, 2 * n, y->opt.max_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_TRY(ensure_classify_buffers(y, width, height));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : y->stream;
  TOD_TRY(classify_device(y, d_frames, n, width, height, d_target, s));
  return order_after_caller(y, s);
}

int tod_yolact_classify_batch(tod_yolact* y, uint32_t* frames, int n, int width, int height) {
  if (!y || !frames) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_classify_batch: null argument");
  if (n < 1 || 2 * n > y->opt.max_tiles) return fail(TOD_ERR_CAPACITY, "classify: %d frames need %d tiles, handle holds %d", n, 2 * n, y->opt.max_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  TOD_TRY(ensure_classify_buffers(y, width, height));
  cudaStream_t s = y->stream;
  const size_t bytes = size_t(n) * width * height * 4;
  TOD_CUDA(cudaMemcpyAsync(y->d_frames, frames, bytes, cudaMemcpyHostToDevice, s));
  TOD_TRY(classify_device(y, y->d_frames, n, width, height, nullptr, s));
  TOD_CUDA(cudaMemcpyAsync(frames, y->d_frames, bytes, cudaMemcpyDeviceToHost, s));  // yolact.rs:233 copy_from_slice
  TOD_TRY(enqueue_diverged(y, 2 * n, s));
  TOD_TRY(wait_stream(y, s, y->done_ev));
  return read_diverged(y, 2 * n);
}

int tod_yolact_classify(tod_yolact* y, uint32_t* frame, int width, int height) { return tod_yolact_classify_batch(y, frame, 1, width, height); }

int tod_yolact_stats(const tod_yolact* y, int64_t* macs_per_tile, int32_t* launches_per_call, int32_t* tc_conv_layers) {
  if (!y) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_stats: null handle");
  if (macs_per_tile) *macs_per_tile = y->macs_per_tile;
  if (launches_per_call) *launches_per_call = y->launches_per_call + (y->seg_out >= 0 ? 1 : 0) + (y->det_ready && y->have_priors ? 5 : 0);
  if (tc_conv_layers) *tc_conv_layers = y->tc_layers;
  return TOD_OK;
}

int tod_yolact_step_macs(const tod_yolact* y, int64_t* macs, int cap) {
  if (!y) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_step_macs: null handle");
  int i = 0;
  for (const Step& st : y->steps) {
    if (macs && i < cap) macs[i] = st.macs;
    ++i;
  }
  return i;
}

// One step enqueued on the lane streams directly (no graph) with a timing event behind every launch: end time of each
// step relative to the start of the step, and the lane it ran on.  Entries [steps .. steps+2] are the literal
// segmentation pass, the box decode / NMS / top-k and the mask assembly (the step's end).
int tod_yolact_trace_steps(tod_yolact* y, int n, float* end_ms, int32_t* lanes, int32_t* kinds, int cap) {
  if (!y) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_trace_steps: null handle");
  if (n < 1 || n > y->opt.max_tiles) return fail(TOD_ERR_CAPACITY, "tod_yolact_trace_steps: n=%d outside [1,%d]", n, y->opt.max_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  const size_t ns = y->steps.size();
  const bool dets = y->det_ready && y->have_priors;
  y->trace_events.assign(ns + 4, nullptr);
  y->trace_lanes.assign(ns, 0);
  for (cudaEvent_t& e : y->trace_events) TOD_CUDA(cudaEventCreate(&e));
  int rc = TOD_OK;
  for (int rep = 0; rep < 3 && rc == TOD_OK; ++rep) {  // the last repetition is the one read back
    rc = enqueue_all_parallel(y, n, dets, dets ? 1 : 0, y->stream);
    if (rc == TOD_OK && cudaStreamSynchronize(y->stream) != cudaSuccess) rc = fail(TOD_ERR_CUDA, "trace run failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  int count = 0;
  if (rc == TOD_OK) {
    for (size_t i = 0; i < ns + 3; ++i) {
      const size_t ev = i < ns ? i : i + 1;
      float t = -1.f;
      if (cudaEventQuery(y->trace_events[ev]) == cudaSuccess) cudaEventElapsedTime(&t, y->trace_events[ns], y->trace_events[ev]);
      (void)cudaGetLastError();
      if (int(i) < cap) {
        if (end_ms) end_ms[i] = t;
        if (lanes) lanes[i] = i < ns ? y->trace_lanes[i] : -1;
        if (kinds) kinds[i] = i < ns ? (y->graph.ops[y->steps[i].op].code | (y->steps[i].kind == kStepConvTc ? 0x1000 : 0) | (y->steps[i].kind == kStepCopy ? 0x2000 : 0)) : -1;
      }
      ++count;
    }
  }
  for (cudaEvent_t e : y->trace_events) cudaEventDestroy(e);
  y->trace_events.clear();
  y->last_tiles = n;
  return rc < 0 ? rc : count;
}

int tod_yolact_profile_ops(tod_yolact* y, int n, float* ms, int32_t* kinds, int cap) {
  if (!y) return fail(TOD_ERR_INVALID_ARG, "tod_yolact_profile_ops: null handle");
  if (n < 1 || n > y->opt.max_tiles) return fail(TOD_ERR_CAPACITY, "tod_yolact_profile_ops: n=%d outside [1,%d]", n, y->opt.max_tiles);
  TOD_CUDA(cudaSetDevice(y->device));
  cudaEvent_t e0, e1;
  TOD_CUDA(cudaEventCreate(&e0));
  TOD_CUDA(cudaEventCreate(&e1));
  cudaStream_t s = y->stream;
  // TOD_PROFILE_REPS > 1: back-to-back launches per event pair, which takes the event/launch gap out of short kernels
  const int reps = std::getenv("TOD_PROFILE_REPS") ? std::max(1, std::atoi(std::getenv("TOD_PROFILE_REPS"))) : 1;
  int i = 0;
  for (const Step& st : y->steps) {
    TOD_TRY(run_step(y, st, n, s));  // warm
    TOD_CUDA(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) TOD_TRY(run_step(y, st, n, s));
    TOD_CUDA(cudaEventRecord(e1, s));
    TOD_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    TOD_CUDA(cudaEventElapsedTime(&t, e0, e1));
    t /= float(reps);
    if (i < cap) {
      if (ms) ms[i] = t;
      if (kinds) kinds[i] = y->graph.ops[st.op].code | (st.kind == kStepConvTc ? 0x1000 : 0) | (st.kind == kStepCopy ? 0x2000 : 0);
    }
    ++i;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  y->last_tiles = n;
  return i;
}

}  // extern "C"
