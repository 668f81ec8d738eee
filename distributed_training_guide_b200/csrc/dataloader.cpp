// Native token-stream loader: the data path of the reference is `datasets` (Arrow, C++) + `tokenizers`
// (Rust) + a Python DataLoader worker process that collates into pageable memory.  For pre-tokenised
// corpora this replaces all of it with ~150 lines of C++: the token file is mmap'ed, a background thread
// cuts it into [batch, seq] int64 batches (shuffled per epoch, partitioned over data-parallel ranks like
// DistributedSampler with drop_last) and writes them into a ring of PINNED host buffers, so the training
// loop's `next()` is a queue pop and the H2D copy is a true async DMA.
#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <torch/extension.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <random>
#include <thread>

#include "comm_api.h"

namespace dtg {
namespace {

class TokenLoader {
 public:
  TokenLoader(const std::string& path, int64_t token_bytes, int64_t seq_len, int64_t batch, int64_t dp_rank,
              int64_t dp_size, int64_t seed, int64_t depth, bool pin)
      : token_bytes_(token_bytes), seq_(seq_len), batch_(batch), rank_(dp_rank), world_(dp_size), seed_(seed),
        depth_(std::max<int64_t>(2, depth)) {
    TORCH_CHECK(token_bytes == 2 || token_bytes == 4, "token file must hold uint16 or uint32 ids");
    fd_ = ::open(path.c_str(), O_RDONLY);
    TORCH_CHECK(fd_ >= 0, "cannot open ", path);
    struct stat st;
    TORCH_CHECK(fstat(fd_, &st) == 0, "cannot stat ", path);
    bytes_ = (size_t)st.st_size;
    base_ = (const uint8_t*)mmap(nullptr, bytes_, PROT_READ, MAP_PRIVATE, fd_, 0);
    TORCH_CHECK(base_ != MAP_FAILED, "mmap failed for ", path);
    madvise((void*)base_, bytes_, MADV_RANDOM);
    n_chunks_ = (int64_t)(bytes_ / token_bytes_) / seq_;
    per_rank_ = n_chunks_ / world_;       // drop_last across ranks
    n_batches_ = per_rank_ / batch_;      // drop_last inside the rank
    TORCH_CHECK(n_batches_ > 0, "token file too small for one batch per rank");
    auto opts = torch::TensorOptions().dtype(torch::kInt64);
    for (int64_t i = 0; i < depth_; ++i) {
      auto t = torch::empty({batch_, seq_}, opts);
      ring_.push_back(pin ? t.pin_memory() : t);
    }
    copied_.assign((size_t)depth_, nullptr);
    guard_.assign((size_t)depth_, 0);
    set_epoch(0);
  }

  ~TokenLoader() {
    stop();
    for (auto e : copied_)
      if (e) cudaEventDestroy(e);
    if (base_ && base_ != MAP_FAILED) munmap((void*)base_, bytes_);
    if (fd_ >= 0) ::close(fd_);
  }

  int64_t num_batches() const { return n_batches_; }
  int64_t num_chunks() const { return n_chunks_; }

  // (re)start the producer for `epoch`: a new permutation, position 0
  void set_epoch(int64_t epoch) {
    stop();
    order_.resize(n_chunks_);
    std::iota(order_.begin(), order_.end(), 0);
    std::mt19937_64 gen((uint64_t)seed_ * 1000003ull + (uint64_t)epoch);
    std::shuffle(order_.begin(), order_.end(), gen);
    head_ = tail_ = 0;
    produced_ = consumed_ = 0;
    done_ = false;
    worker_ = std::thread([this] { this->run(); });
  }

  // next batch: a view of a pinned ring slot.  The slot becomes refillable `depth - 1` calls later AND, if the
  // consumer reported an asynchronous copy out of it (`mark_copied`), only once that copy has executed.
  torch::Tensor next() {
    std::unique_lock<std::mutex> lk(mu_);
    TORCH_CHECK(consumed_ < n_batches_, "epoch exhausted: call set_epoch()");
    cv_.wait(lk, [this] { return produced_ > consumed_; });
    last_slot_ = consumed_ % depth_;
    torch::Tensor out = ring_[last_slot_];
    ++consumed_;
    cv_.notify_all();
    return out;
  }

  // The consumer enqueued `non_blocking` H2D copies out of the slot handed out by the last next() on the CURRENT
  // CUDA stream: record an event behind them; the producer waits on it before overwriting the slot.  (A call
  // count alone is not enough: the host may run many steps ahead of the device, so the DMA that reads the slot
  // may not have executed when the ring wraps — torch's pinned caching allocator guards this with stream events
  // too.)
  void mark_copied() {
    int64_t slot;
    {
      std::lock_guard<std::mutex> lk(mu_);
      slot = last_slot_;
    }
    if (slot < 0) return;
    cudaEvent_t& e = copied_[(size_t)slot];
    if (!e) C10_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    C10_CUDA_CHECK(cudaEventRecord(e, at::cuda::getCurrentCUDAStream().stream()));
    std::lock_guard<std::mutex> lk(mu_);
    guard_[(size_t)slot] = 1;
  }

 private:
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      done_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
  }

  void run() {
    for (int64_t b = 0; b < n_batches_; ++b) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        // keep one slot of slack: the consumer may still be copying the slot it was handed last
        cv_.wait(lk, [this] { return done_ || produced_ - consumed_ < depth_ - 1; });
        if (done_) return;
      }
      {
        cudaEvent_t e = nullptr;
        {
          std::lock_guard<std::mutex> lk(mu_);
          if (guard_[(size_t)(b % depth_)]) {
            e = copied_[(size_t)(b % depth_)];
            guard_[(size_t)(b % depth_)] = 0;
          }
        }
        if (e) cudaEventSynchronize(e);  // the DMA out of this slot has executed
      }
      int64_t* dst = ring_[b % depth_].data_ptr<int64_t>();
      for (int64_t i = 0; i < batch_; ++i) {
        // DistributedSampler layout: rank r takes order[r], order[r + world], ...
        const int64_t chunk = order_[(b * batch_ + i) * world_ + rank_];
        const uint8_t* src = base_ + (size_t)chunk * seq_ * token_bytes_;
        if (token_bytes_ == 2) {
          const uint16_t* s = (const uint16_t*)src;
          for (int64_t k = 0; k < seq_; ++k) dst[i * seq_ + k] = s[k];
        } else {
          const uint32_t* s = (const uint32_t*)src;
          for (int64_t k = 0; k < seq_; ++k) dst[i * seq_ + k] = s[k];
        }
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        ++produced_;
      }
      cv_.notify_all();
    }
  }

  int64_t token_bytes_, seq_, batch_, rank_, world_, seed_, depth_;
  int fd_ = -1;
  const uint8_t* base_ = nullptr;
  size_t bytes_ = 0;
  int64_t n_chunks_ = 0, per_rank_ = 0, n_batches_ = 0;
  std::vector<int64_t> order_;
  std::vector<torch::Tensor> ring_;
  std::vector<cudaEvent_t> copied_;  // per slot: recorded behind the consumer's async copies out of it
  std::vector<char> guard_;          // per slot: copied_[slot] must be waited for before the next refill
  int64_t last_slot_ = -1;
  std::thread worker_;
  std::mutex mu_;
  std::condition_variable cv_;
  int64_t head_ = 0, tail_ = 0, produced_ = 0, consumed_ = 0;
  bool done_ = false;
};

}  // namespace

void bind_dataloader(pybind11::module_& m) {
  pybind11::class_<TokenLoader>(m, "TokenLoader")
      .def(pybind11::init<const std::string&, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, bool>(),
           pybind11::arg("path"), pybind11::arg("token_bytes"), pybind11::arg("seq_len"), pybind11::arg("batch"),
           pybind11::arg("dp_rank") = 0, pybind11::arg("dp_size") = 1, pybind11::arg("seed") = 0,
           pybind11::arg("depth") = 4, pybind11::arg("pin") = true)
      .def("num_batches", &TokenLoader::num_batches)
      .def("num_chunks", &TokenLoader::num_chunks)
      .def("set_epoch", &TokenLoader::set_epoch, pybind11::call_guard<pybind11::gil_scoped_release>())
      .def("next", &TokenLoader::next, pybind11::call_guard<pybind11::gil_scoped_release>())
      .def("mark_copied", &TokenLoader::mark_copied);
}
}  // namespace dtg
