// Frame sharder behind the C ABI: independent camera frames / tiles are cut into contiguous ranges,
// one range per GPU, each range into chunks that the GPU's handles take in turn - one host thread per handle, `depth`
// handles per GPU, so one chunk's host<->device copies and latency-bound tail overlap the next chunk's backbone (what
// the frame loop's double buffering does), and results land in the caller's buffers at the frame's index.  There is no
// collective and no cross-GPU traffic: the path needs none.  Every chunk runs through the ordinary single-handle
// entry points (tod_yolact_infer_tiles_cells, tod_yolact_classify_batch, tod_scene_append_batch_device), so the bytes
// of a frame do not depend on how many GPUs or handles the pool has.
//
// Replaces, for a multi-GPU box, the reference's one-frame-at-a-time loop (`process_scene` -> `Yolact::classify`,
// src/scene.rs:77-119; `manage` -> `append_scene`, src/main.rs:78-96).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

namespace tod {
namespace {

struct Worker;
using Job = std::function<int(Worker&)>;

struct DeviceQueue {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Job> jobs;
  bool stop = false;
};

struct Worker {
  int device = 0;
  tod_yolact* y = nullptr;
  // fused RGB-D scratch (allocated on first use for a given frame size)
  tod_scene* scene = nullptr;
  int sw = 0, sh = 0, scap = 0;
  uint32_t *d_frames = nullptr, *d_map = nullptr;
  uint16_t *d_depth = nullptr, *d_target = nullptr;
  float *d_world = nullptr, *d_conn0 = nullptr, *d_conn1 = nullptr, *d_balls = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  std::thread thread;
  DeviceQueue* q = nullptr;
};

}  // namespace
}  // namespace tod

using namespace tod;

struct tod_pool {
  std::vector<int> devices;
  int depth = 1, max_tiles = 0;
  std::vector<DeviceQueue*> queues;   // one per device
  std::vector<Worker*> workers;       // devices x depth
  // completion of the current call
  std::mutex mu;
  std::condition_variable cv;
  int pending = 0, first_error = 0, warn = 0;
  std::string error_text;
  // model facts needed to slice the caller's buffers
  int n_outputs = 0, tile_w = 0, tile_h = 0, gh = 0, gw = 0, ph = 0, pw = 0, max_dets = 0;
  std::vector<int64_t> out_elems;
};

namespace {

void worker_main(tod_pool* p, Worker* w) {
  cudaSetDevice(w->device);
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(w->q->mu);
      w->q->cv.wait(lk, [&] { return w->q->stop || !w->q->jobs.empty(); });
      if (w->q->jobs.empty()) return;  // stop
      job = std::move(w->q->jobs.front());
      w->q->jobs.pop_front();
    }
    const int rc = job(*w);
    std::lock_guard<std::mutex> lk(p->mu);
    if (rc < 0 && p->first_error == 0) {
      p->first_error = rc;
      p->error_text = last_error();   // the worker's thread-local text, handed to the caller's thread
    }
    if (rc > 0) p->warn = rc;
    if (--p->pending == 0) p->cv.notify_all();
  }
}

// runs jobs[d] on device d's handles and blocks until all are done; returns the first error, else a warning code, else 0
int run_jobs(tod_pool* p, std::vector<std::vector<Job>>& jobs) {
  int total = 0;
  for (auto& v : jobs) total += int(v.size());
  if (total == 0) return TOD_OK;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->pending = total;
    p->first_error = 0;
    p->warn = 0;
  }
  for (size_t d = 0; d < jobs.size(); ++d) {
    {
      std::lock_guard<std::mutex> lk(p->queues[d]->mu);
      for (Job& j : jobs[d]) p->queues[d]->jobs.push_back(std::move(j));
    }
    p->queues[d]->cv.notify_all();
  }
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv.wait(lk, [&] { return p->pending == 0; });
  if (p->first_error < 0) return fail(p->first_error, "%s", p->error_text.c_str());
  return p->warn;
}

// contiguous shard of [0, n) for device d of G (ceil(n / G) units each)
inline void shard(int n, int G, int d, int* lo, int* hi) {
  const int per = (n + G - 1) / G;
  *lo = std::min(n, d * per);
  *hi = std::min(n, *lo + per);
}

void free_rgbd(Worker& w) {
  if (w.scene) tod_scene_destroy(w.scene);
  w.scene = nullptr;
  for (void* q : {(void*)w.d_frames, (void*)w.d_map, (void*)w.d_depth, (void*)w.d_target, (void*)w.d_world, (void*)w.d_conn0, (void*)w.d_conn1, (void*)w.d_balls})
    if (q) cudaFree(q);
  w.d_frames = w.d_map = nullptr;
  w.d_depth = w.d_target = nullptr;
  w.d_world = w.d_conn0 = w.d_conn1 = w.d_balls = nullptr;
  w.sw = w.sh = w.scap = 0;
}

int ensure_rgbd(Worker& w, const tod_scene_params& sp, int cap) {
  if (w.scene && w.sw == sp.width && w.sh == sp.height && w.scap == cap) return TOD_OK;
  free_rgbd(w);
  tod_scene_params p = sp;
  p.max_batch = cap;
  TOD_TRY(tod_scene_create(w.device, &p, &w.scene));
  const size_t npx = size_t(sp.width) * sp.height, nb = size_t(cap);
  TOD_CUDA(cudaMalloc(&w.d_frames, nb * npx * 4));
  TOD_CUDA(cudaMalloc(&w.d_depth, nb * npx * 2));
  TOD_CUDA(cudaMalloc(&w.d_target, nb * npx * 2));
  TOD_CUDA(cudaMalloc(&w.d_map, nb * npx * 4));
  TOD_CUDA(cudaMalloc(&w.d_world, nb * npx * 16));
  TOD_CUDA(cudaMalloc(&w.d_conn0, nb * npx * 16));
  TOD_CUDA(cudaMalloc(&w.d_conn1, nb * npx * 16));
  TOD_CUDA(cudaMalloc(&w.d_balls, nb * 100 * 16));
  w.sw = sp.width;
  w.sh = sp.height;
  w.scap = cap;
  return TOD_OK;
}

}  // namespace

extern "C" {

void tod_pool_destroy(tod_pool* p) {
  if (!p) return;
  for (DeviceQueue* q : p->queues) {
    {
      std::lock_guard<std::mutex> lk(q->mu);
      q->stop = true;
    }
    q->cv.notify_all();
  }
  for (Worker* w : p->workers) {
    if (w->thread.joinable()) w->thread.join();
    cudaSetDevice(w->device);
    free_rgbd(*w);
    if (w->done) cudaEventDestroy(w->done);
    if (w->stream) cudaStreamDestroy(w->stream);
    if (w->y) tod_yolact_destroy(w->y);
    delete w;
  }
  for (DeviceQueue* q : p->queues) delete q;
  delete p;
}

int tod_pool_create(const char* tflite_path, const int32_t* devices, int n_devices, int depth, const tod_yolact_options* opts, tod_pool** out) {
  if (!tflite_path || !out || (n_devices > 0 && !devices)) return fail(TOD_ERR_INVALID_ARG, "tod_pool_create: null argument");
  *out = nullptr;
  int avail = 0;
  TOD_TRY(tod_device_count(&avail));
  if (avail < 1) return fail(TOD_ERR_NO_DEVICE, "tod_pool_create: no sm_100 device is visible");
  if (depth < 1 || depth > 8) return fail(TOD_ERR_INVALID_ARG, "tod_pool_create: depth must be in [1,8]");
  tod_yolact_options o;
  tod_yolact_default_options(&o);
  if (opts) o = *opts;
  if (!opts || o.batches_in_flight < depth) o.batches_in_flight = depth;  // the pool's handles share their GPU
  tod_pool* p = new tod_pool();
  if (n_devices <= 0)   // every visible device
    for (int d = 0; d < avail; ++d) p->devices.push_back(d);
  else
    p->devices.assign(devices, devices + n_devices);
  p->depth = depth;
  p->max_tiles = o.max_tiles;
  p->max_dets = o.max_dets;
  for (size_t d = 0; d < p->devices.size(); ++d) p->queues.push_back(new DeviceQueue());
  for (size_t d = 0; d < p->devices.size(); ++d)
    for (int k = 0; k < depth; ++k) {
      Worker* w = new Worker();
      w->device = p->devices[d];
      w->q = p->queues[d];
      p->workers.push_back(w);
      int rc = tod_yolact_create(tflite_path, w->device, &o, &w->y);
      if (rc == TOD_OK && cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(TOD_ERR_CUDA, "tod_pool_create: cudaStreamCreate failed");
      if (rc == TOD_OK && cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) rc = fail(TOD_ERR_CUDA, "tod_pool_create: cudaEventCreate failed");
      if (rc < 0) {
        tod_pool_destroy(p);
        return rc;
      }
    }
  tod_yolact* y0 = p->workers[0]->y;
  p->n_outputs = tod_yolact_num_outputs(y0);
  for (int k = 0; k < p->n_outputs; ++k) {
    int32_t shp[4], elems = 0;
    tod_yolact_output_info(y0, k, shp, nullptr, nullptr, &elems);
    p->out_elems.push_back(elems);
    if (shp[1] > 1 && k == 4) { p->gh = shp[1]; p->gw = shp[2]; }
    if (shp[1] > 1 && k != 4) { p->ph = shp[1]; p->pw = shp[2]; }
  }
  {
    int32_t shp[4];
    // the model input is tensor `inputs[0]`; its shape is what a tile is (yolact.rs:143-145)
    tod_yolact_input_info(y0, shp);
    p->tile_h = shp[1];
    p->tile_w = shp[2];
  }
  for (Worker* w : p->workers) w->thread = std::thread(worker_main, p, w);
  *out = p;
  return TOD_OK;
}

int tod_pool_num_handles(const tod_pool* p) { return p ? int(p->workers.size()) : 0; }
int tod_pool_num_devices(const tod_pool* p) { return p ? int(p->devices.size()) : 0; }

int tod_pool_infer_tiles(tod_pool* p, const uint8_t* rgb_tiles, int n, uint8_t* const* outputs_u8, uint32_t* tile_classes,
                         uint32_t* cell_classes, tod_detections* dets) {
  if (!p || !rgb_tiles) return fail(TOD_ERR_INVALID_ARG, "tod_pool_infer_tiles: null argument");
  if (n < 1) return fail(TOD_ERR_INVALID_ARG, "tod_pool_infer_tiles: n must be positive");
  if (dets && dets->max_dets != p->max_dets) return fail(TOD_ERR_INVALID_ARG, "tod_detections.max_dets=%d, pool was created with %d", dets->max_dets, p->max_dets);
  const int G = int(p->devices.size());
  const size_t tile_bytes = size_t(p->tile_w) * p->tile_h * 3;
  std::vector<std::vector<Job>> jobs(G);
  for (int d = 0; d < G; ++d) {
    int lo, hi;
    shard(n, G, d, &lo, &hi);
    for (int c0 = lo; c0 < hi; c0 += p->max_tiles) {
      const int cn = std::min(p->max_tiles, hi - c0);
      jobs[d].push_back([=](Worker& w) -> int {
        std::vector<uint8_t*> outs;
        if (outputs_u8) {
          outs.assign(p->n_outputs, nullptr);
          for (int k = 0; k < p->n_outputs; ++k)
            if (outputs_u8[k]) outs[k] = outputs_u8[k] + size_t(c0) * p->out_elems[k];
        }
        tod_detections dd{};
        if (dets) {
          const size_t md = size_t(dets->max_dets), t0 = size_t(c0), px = size_t(p->ph) * p->pw;
          dd.max_dets = dets->max_dets;
          dd.count = dets->count ? dets->count + t0 : nullptr;
          dd.boxes = dets->boxes ? dets->boxes + t0 * md * 4 : nullptr;
          dd.scores = dets->scores ? dets->scores + t0 * md : nullptr;
          dd.classes = dets->classes ? dets->classes + t0 * md : nullptr;
          dd.priors = dets->priors ? dets->priors + t0 * md : nullptr;
          dd.masks = dets->masks ? dets->masks + t0 * md * px : nullptr;
          dd.masks_bin = dets->masks_bin ? dets->masks_bin + t0 * md * px : nullptr;
          dd.masks_bits = dets->masks_bits ? dets->masks_bits + t0 * md * ((px + 31) / 32) : nullptr;
          dd.masks_tile_bits = dets->masks_tile_bits ? dets->masks_tile_bits + t0 * md * ((size_t(p->tile_w) * p->tile_h + 31) / 32) : nullptr;
        }
        return tod_yolact_infer_tiles_cells(w.y, rgb_tiles + size_t(c0) * tile_bytes, cn, outputs_u8 ? outs.data() : nullptr,
                                            tile_classes ? tile_classes + size_t(c0) * p->tile_w * p->tile_h : nullptr,
                                            cell_classes ? cell_classes + size_t(c0) * p->gh * p->gw : nullptr, dets ? &dd : nullptr);
      });
    }
  }
  return run_jobs(p, jobs);
}

int tod_pool_classify_batch(tod_pool* p, uint32_t* frames, int n, int width, int height) {
  if (!p || !frames) return fail(TOD_ERR_INVALID_ARG, "tod_pool_classify_batch: null argument");
  if (n < 1 || width < 2 || height < 2) return fail(TOD_ERR_INVALID_ARG, "tod_pool_classify_batch: bad size");
  const int G = int(p->devices.size());
  const int cap = std::max(1, p->max_tiles / 2);   // frames per chunk: two tiles per frame (yolact.rs:213-217)
  if (p->max_tiles < 2) return fail(TOD_ERR_CAPACITY, "tod_pool_classify_batch: handles hold %d tile(s), a frame needs two", p->max_tiles);
  const size_t npx = size_t(width) * height;
  std::vector<std::vector<Job>> jobs(G);
  for (int d = 0; d < G; ++d) {
    int lo, hi;
    shard(n, G, d, &lo, &hi);
    for (int c0 = lo; c0 < hi; c0 += cap) {
      const int cn = std::min(cap, hi - c0);
      jobs[d].push_back([=](Worker& w) -> int { return tod_yolact_classify_batch(w.y, frames + size_t(c0) * npx, cn, width, height); });
    }
  }
  return run_jobs(p, jobs);
}

// The fused RGB-D frame loop (scene.rs:84-97 feeding scene.rs:147-331) for n frames: classify in place -> target =
// low 16 bits (scene.rs:93) -> pt_cloud + pt_cloud_weights, with the target never leaving the GPU.
int tod_pool_rgbd_batch(tod_pool* p, const tod_scene_params* sp, uint32_t* frames, const uint16_t* depth, int n, uint32_t* map,
                        float* world4, float* conn0, float* conn1, float* balls4) {
  if (!p || !sp || !frames || !depth) return fail(TOD_ERR_INVALID_ARG, "tod_pool_rgbd_batch: null argument");
  if (n < 1) return fail(TOD_ERR_INVALID_ARG, "tod_pool_rgbd_batch: n must be positive");
  if (p->max_tiles < 2) return fail(TOD_ERR_CAPACITY, "tod_pool_rgbd_batch: handles hold %d tile(s), a frame needs two", p->max_tiles);
  const int G = int(p->devices.size());
  const int cap = p->max_tiles / 2;
  const tod_scene_params prm = *sp;
  const size_t npx = size_t(sp->width) * sp->height;
  std::vector<std::vector<Job>> jobs(G);
  for (int d = 0; d < G; ++d) {
    int lo, hi;
    shard(n, G, d, &lo, &hi);
    for (int c0 = lo; c0 < hi; c0 += cap) {
      const int cn = std::min(cap, hi - c0);
      jobs[d].push_back([=](Worker& w) -> int {
        TOD_TRY(ensure_rgbd(w, prm, cap));
        cudaStream_t s = w.stream;
        const size_t o = size_t(c0) * npx, m = size_t(cn) * npx;
        TOD_CUDA(cudaMemcpyAsync(w.d_frames, frames + o, m * 4, cudaMemcpyHostToDevice, s));
        TOD_CUDA(cudaMemcpyAsync(w.d_depth, depth + o, m * 2, cudaMemcpyHostToDevice, s));
        TOD_TRY(tod_yolact_classify_batch_device(w.y, w.d_frames, cn, prm.width, prm.height, w.d_target, s));
        TOD_TRY(tod_scene_append_batch_device(w.scene, w.d_depth, w.d_target, cn, w.d_map, world4 ? w.d_world : nullptr, conn0 ? w.d_conn0 : nullptr,
                                              conn1 ? w.d_conn1 : nullptr, balls4 ? w.d_balls : nullptr, s));
        TOD_CUDA(cudaMemcpyAsync(frames + o, w.d_frames, m * 4, cudaMemcpyDeviceToHost, s));   // yolact.rs:233: classify mutates the frame
        if (map) TOD_CUDA(cudaMemcpyAsync(map + o, w.d_map, m * 4, cudaMemcpyDeviceToHost, s));
        if (world4) TOD_CUDA(cudaMemcpyAsync(world4 + o * 4, w.d_world, m * 16, cudaMemcpyDeviceToHost, s));
        if (conn0) TOD_CUDA(cudaMemcpyAsync(conn0 + o * 4, w.d_conn0, m * 16, cudaMemcpyDeviceToHost, s));
        if (conn1) TOD_CUDA(cudaMemcpyAsync(conn1 + o * 4, w.d_conn1, m * 16, cudaMemcpyDeviceToHost, s));
        if (balls4) TOD_CUDA(cudaMemcpyAsync(balls4 + size_t(c0) * 400, w.d_balls, size_t(cn) * 1600, cudaMemcpyDeviceToHost, s));
        TOD_CUDA(cudaEventRecord(w.done, s));
        TOD_CUDA(cudaEventSynchronize(w.done));
        int div = 0;
        TOD_TRY(tod_yolact_last_diverged(w.y, &div));
        return div ? TOD_WARN_REFERENCE_DIVERGES : TOD_OK;
      });
    }
  }
  return run_jobs(p, jobs);
}

}  // extern "C"
