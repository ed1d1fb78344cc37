// host_copy.h -- a small persistent thread pool that copies host memory in parallel.
//
// Why it exists: vierkant hands vierkant::bcn::compress() an image decoded into malloc'ed memory and receives its blocks
// in std::vector storage (include/vierkant/texture_block_compression.hpp:27-44) -- pageable memory on both sides.  CUDA
// copies from / to pageable memory are staged by the driver on the calling thread and block it, which serialises the
// band pipeline of compress() (measured: 4096^2 chain 3.3 ms with pinned buffers, 16.0 ms with pageable ones).  The
// library therefore stages pageable buffers itself, through pinned buffers it owns, and moves the bytes between the
// caller's memory and those with several threads while the GPU works on the bands already queued.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace vkt
{

// memcpy whose stores bypass the cache (MOVNTDQ): the destinations here are read next by a DMA engine (pinned staging) or
// much later by the caller (its block vectors), never by this core, so write-allocating their lines only adds a third
// stream of memory traffic to a copy that is bandwidth-bound (measured on the B200 box's host: 19 -> 2x GB/s with 8 threads).
static inline void stream_copy(void *dst, const void *src, size_t bytes)
{
#if defined(__SSE2__)
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    if(bytes < 4096)
    {
        std::memcpy(d, s, bytes);
        return;
    }
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    std::memcpy(d, s, head);
    d += head, s += head, bytes -= head;
    size_t n64 = bytes / 64;
    for(; n64; --n64, d += 64, s += 64)
    {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d), a), _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c), _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
    }
    _mm_sfence();
    std::memcpy(d, s, bytes & 63);
#else
    std::memcpy(dst, src, bytes);
#endif
}

class CopyPool
{
public:
    explicit CopyPool(unsigned workers)
    {
        for(unsigned i = 0; i < workers; ++i) { threads_.emplace_back([this] { run(); }); }
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        work_.notify_all();
        for(auto &t: threads_) { t.join(); }
    }
    CopyPool(const CopyPool &) = delete;
    CopyPool &operator=(const CopyPool &) = delete;

    // memcpy(dst, src, bytes), split over the workers and the calling thread; returns when every byte is in place
    void copy(void *dst, const void *src, size_t bytes)
    {
        constexpr size_t kMinChunk = size_t(256) << 10;
        const size_t parts = std::min<size_t>(threads_.size() + 1, bytes / kMinChunk);
        if(parts <= 1)
        {
            stream_copy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> one_call(call_);// one parallel copy at a time (callers of different slots may meet here)
        const size_t chunk = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095);
        size_t mine = 0;
        {
            std::lock_guard<std::mutex> g(m_);
            for(size_t off = 0; off < bytes; off += chunk)
            {
                const size_t n = std::min(chunk, bytes - off);
                if(off == 0) { mine = n; }
                else { jobs_.push_back({static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, n}), ++outstanding_; }
            }
        }
        work_.notify_all();
        stream_copy(dst, src, mine);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return outstanding_ == 0; });
    }

private:
    struct Job
    {
        char *dst;
        const char *src;
        size_t bytes;
    };
    void run()
    {
        std::unique_lock<std::mutex> g(m_);
        for(;;)
        {
            work_.wait(g, [this] { return stop_ || !jobs_.empty(); });
            if(jobs_.empty())
            {
                if(stop_) { return; }
                continue;
            }
            const Job j = jobs_.front();
            jobs_.pop_front();
            g.unlock();
            stream_copy(j.dst, j.src, j.bytes);
            g.lock();
            if(--outstanding_ == 0) { done_.notify_all(); }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_, call_;
    std::condition_variable work_, done_;
    std::deque<Job> jobs_;
    size_t outstanding_ = 0;
    bool stop_ = false;
};

}// namespace vkt
