// Frame pool: one context and one host worker thread per GPU inside one process ("frame f -> GPU
// f*G/N, one host worker thread + >= 2 streams per GPU, pinned double-buffered staging").  Pure host code over the
// C ABI of context.cu -- the reference has no counterpart, its Compute() takes one frame on one GPU
// (InstanceStixels/src/Stixels.cu:449-637, frame loop apps/run_cityscapes.cu:249-449).
//
// A call shards its n frames into contiguous blocks, worker w takes frames [n*w/G, n*(w+1)/G) and streams them
// through isx_submit_batch_host / isx_wait_batch_host in sub-batches of at most `max_batch` frames, three in flight.
// Frames are independent, so there is no collective: every worker writes its frames' Sections into the caller's
// array at their place; the instance records are concatenated in frame order when all workers are done.
#include <condition_variable>
#include <memory>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/instance_stixels_b200.h"

namespace {

struct Job {
  int pairwise = 0, first = 0, count = 0;
  const float *disparity = nullptr;
  const int32_t *segmentation = nullptr;
  const isx_road *roads = nullptr;
  isx_section *sections = nullptr;
  bool want_instances = false;
};

struct Worker {
  isx_handle h = nullptr;
  int device = 0;
  std::thread thread;
  std::mutex m;
  std::condition_variable cv;
  bool has_job = false, done = false, quit = false;
  Job job;
  int rc = ISX_OK;
  std::string error;
  // results of the last job
  std::vector<isx_instance> inst;
  std::vector<int32_t> inst_count;  // per frame of the block
  std::unique_ptr<isx_instance[]> tmp;
  size_t tmp_cap = 0;
};

}  // namespace

struct isx_pool {
  std::vector<Worker *> workers;
  int max_batch = 0;
  size_t hw = 0, seg_elems = 0, sec_per_frame = 0;
  int realcols = 0, inst_cap = 0;
  std::string last_error;
};

namespace {

thread_local std::string g_pool_error;

int pool_fail(isx_pool *p, int code, const std::string &msg) {
  g_pool_error = msg;
  if (p) p->last_error = msg;
  return code;
}

// One block of frames through one context: sub-batches of <= max_batch frames, three in flight.
void run_job(isx_pool *p, Worker *w) {
  const Job &j = w->job;
  w->rc = ISX_OK;
  w->error.clear();
  w->inst.clear();
  w->inst_count.assign((size_t)j.count, 0);
  const int mb = p->max_batch;
  const int nsub = (j.count + mb - 1) / mb;
  // scratch for the records of one sub-batch: kept across calls, never zero-filled (52 MB at 64 frames)
  const size_t tmp_need = j.want_instances ? (size_t)mb * p->inst_cap : 0;
  if (tmp_need > w->tmp_cap) {
    w->tmp.reset(new isx_instance[tmp_need]);
    w->tmp_cap = tmp_need;
  }
  isx_instance *tmp = w->tmp.get();
  std::vector<int32_t> offs((size_t)mb + 1);
  int submitted = 0, waited = 0;
  auto sub_first = [&](int s) { return s * mb; };
  auto sub_count = [&](int s) { return (j.count - s * mb) < mb ? (j.count - s * mb) : mb; };
  while (waited < nsub) {
    while (submitted < nsub && submitted - waited < 3 && w->rc == ISX_OK) {
      const int f0 = j.first + sub_first(submitted), cn = sub_count(submitted);
      const int rc = isx_submit_batch_host(w->h, j.pairwise, cn, j.disparity + (size_t)f0 * p->hw,
                                           j.segmentation + (size_t)f0 * p->seg_elems, j.roads + f0,
                                           j.sections ? j.sections + (size_t)f0 * p->sec_per_frame : nullptr);
      if (rc != ISX_OK) {
        w->rc = rc;
        w->error = isx_last_error(w->h);
        break;
      }
      submitted++;
    }
    if (waited >= submitted) break;  // a submit failed and nothing is in flight any more
    const int cn = sub_count(waited);
    const int rc = isx_wait_batch_host(w->h, j.want_instances ? tmp : nullptr, (int)tmp_need,
                                       j.want_instances ? offs.data() : nullptr);
    if (rc != ISX_OK && w->rc == ISX_OK) {
      w->rc = rc;
      w->error = isx_last_error(w->h);
    }
    if (j.want_instances && rc == ISX_OK) {
      w->inst.insert(w->inst.end(), tmp, tmp + offs[cn]);
      for (int f = 0; f < cn; f++) w->inst_count[(size_t)sub_first(waited) + f] = offs[f + 1] - offs[f];
    }
    waited++;
  }
}

void worker_main(isx_pool *p, Worker *w) {
  std::unique_lock<std::mutex> lk(w->m);
  while (true) {
    w->cv.wait(lk, [&] { return w->has_job || w->quit; });
    if (w->quit) return;
    lk.unlock();
    run_job(p, w);
    lk.lock();
    w->has_job = false;
    w->done = true;
    w->cv.notify_all();
  }
}

}  // namespace

extern "C" {

int isx_pool_create(isx_pool_handle *out, const int *devices, int n_devices, const isx_config *cfg, int max_batch) {
  if (!out || !devices || n_devices < 1 || !cfg || max_batch < 1)
    return pool_fail(nullptr, ISX_ERR_INVALID_ARGUMENT, "isx_pool_create: bad argument");
  isx_pool *p = new isx_pool();
  p->max_batch = max_batch;
  for (int i = 0; i < n_devices; i++) {
    Worker *w = new Worker();
    w->device = devices[i];
    int rc = isx_create(&w->h, devices[i]);
    if (rc == ISX_OK) rc = isx_set_config(w->h, cfg);
    if (rc == ISX_OK) rc = isx_initialize(w->h, max_batch);
    if (rc == ISX_OK) rc = isx_reserve_in_flight(w->h, 3);   // run_job keeps three sub-batches in flight
    if (rc != ISX_OK) {
      const std::string msg = std::string("isx_pool_create, device ") + std::to_string(devices[i]) + ": " +
                              isx_last_error(w->h ? w->h : nullptr);
      if (w->h) isx_destroy(w->h);
      delete w;
      for (Worker *v : p->workers) { isx_destroy(v->h); delete v; }
      delete p;
      return pool_fail(nullptr, rc, msg);
    }
    p->workers.push_back(w);
  }
  isx_handle h0 = p->workers[0]->h;
  p->realcols = isx_real_cols(h0);
  p->seg_elems = isx_segmentation_elems(h0);
  p->hw = (size_t)cfg->rows * (size_t)cfg->cols;
  p->sec_per_frame = (size_t)p->realcols * isx_max_sections(h0);
  p->inst_cap = isx_instance_capacity(h0);
  for (Worker *w : p->workers) w->thread = std::thread(worker_main, p, w);
  *out = p;
  return ISX_OK;
}

int isx_pool_destroy(isx_pool_handle p) {
  if (!p) return ISX_OK;
  for (Worker *w : p->workers) {
    {
      std::lock_guard<std::mutex> lk(w->m);
      w->quit = true;
    }
    w->cv.notify_all();
    if (w->thread.joinable()) w->thread.join();
    isx_destroy(w->h);
    delete w;
  }
  delete p;
  return ISX_OK;
}

int isx_pool_size(isx_pool_handle p) { return p ? (int)p->workers.size() : 0; }
int isx_pool_real_cols(isx_pool_handle p) { return p ? p->realcols : 0; }
size_t isx_pool_segmentation_elems(isx_pool_handle p) { return p ? p->seg_elems : 0; }
const char *isx_pool_last_error(isx_pool_handle p) { return p ? p->last_error.c_str() : g_pool_error.c_str(); }

int isx_pool_compute_host(isx_pool_handle p, int pairwise, int n, const float *disparity,
                          const int32_t *segmentation, const isx_road *roads, isx_section *sections,
                          isx_instance *instances, int instances_capacity, int32_t *instance_offsets) {
  if (!p) return pool_fail(nullptr, ISX_ERR_INVALID_ARGUMENT, "null pool");
  if (n < 1 || !disparity || !segmentation || !roads) return pool_fail(p, ISX_ERR_INVALID_ARGUMENT, "bad argument");
  const int G = (int)p->workers.size();
  const bool want = instances != nullptr || instance_offsets != nullptr;
  // frame f -> worker f * G / n: contiguous blocks
  for (int g = 0; g < G; g++) {
    Worker *w = p->workers[g];
    const int first = (int)((long long)n * g / G), last = (int)((long long)n * (g + 1) / G);
    std::lock_guard<std::mutex> lk(w->m);
    w->job = Job{pairwise, first, last - first, disparity, segmentation, roads, sections, want};
    w->done = false;
    w->has_job = last > first;
    if (!w->has_job) {
      w->done = true;
      w->rc = ISX_OK;
      w->inst.clear();
      w->inst_count.clear();
    }
    w->cv.notify_all();
  }
  int rc = ISX_OK;
  for (int g = 0; g < G; g++) {
    Worker *w = p->workers[g];
    std::unique_lock<std::mutex> lk(w->m);
    w->cv.wait(lk, [&] { return w->done; });
    if (w->rc != ISX_OK && rc == ISX_OK)
      rc = pool_fail(p, w->rc, "worker " + std::to_string(g) + " (device " + std::to_string(w->device) + "): " + w->error);
  }
  if (rc != ISX_OK) return rc;
  if (want) {
    int total = 0, f = 0;
    for (int g = 0; g < G; g++) {
      Worker *w = p->workers[g];
      if (instances) {
        const int room = instances_capacity - total;
        const int take = (int)w->inst.size() < room ? (int)w->inst.size() : (room > 0 ? room : 0);
        if (take > 0) std::memcpy(instances + total, w->inst.data(), sizeof(isx_instance) * (size_t)take);
      }
      for (int32_t c : w->inst_count) {
        if (instance_offsets) instance_offsets[f] = total;
        total += c;
        f++;
      }
    }
    if (instance_offsets) instance_offsets[n] = total;
  }
  return ISX_OK;
}

}  // extern "C"
