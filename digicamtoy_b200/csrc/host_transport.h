// Host side of the 12-bit transport of dct_generate_host: the device writes a chunk's samples as a
// little-endian stream of 12-bit values (store_run_packed12 in generate.cuh: sample s in bits
// [12 s, 12 s + 12)), the copy engine moves 1.5 instead of 2 bytes per sample over PCIe, and a small pool
// of host threads widens them into the caller's int16 buffer.  Plain C++ (no CUDA in here).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace dct {

// n samples from the packed stream `src` (which starts on a sample-pair boundary) into dst
inline void unpack12_scalar(const unsigned char* src, int16_t* dst, size_t n) {
    size_t k = 0;
    for (; k + 1 < n; k += 2, src += 3) {
        dst[k] = (int16_t)(src[0] | ((src[1] & 0x0f) << 8));
        dst[k + 1] = (int16_t)((src[1] >> 4) | (src[2] << 4));
    }
    if (k < n) dst[k] = (int16_t)(src[0] | ((src[1] & 0x0f) << 8));
}

#if defined(__x86_64__)
// 16 samples (24 bytes) per step: two 12-byte groups into the two 128-bit lanes, bytes spread to
// 16-bit lanes by PSHUFB (even samples: bytes 3k, 3k+1 -> & 0xfff; odd: bytes 3k+1, 3k+2 -> >> 4)
__attribute__((target("avx2"))) inline void unpack12_avx2(const unsigned char* src, int16_t* dst, size_t n) {
    const __m256i shuf = _mm256_setr_epi8(0, 1, 1, 2, 3, 4, 4, 5, 6, 7, 7, 8, 9, 10, 10, 11,
                                          0, 1, 1, 2, 3, 4, 4, 5, 6, 7, 7, 8, 9, 10, 10, 11);
    const __m256i mask = _mm256_set1_epi16(0x0fff);
    size_t k = 0;
    // The destination is written once and not read back here: streaming stores (no read-for-ownership
    // of the caller's buffer) once it is 32-byte aligned -- reachable in whole sample pairs only if
    // dst is 4-byte aligned.
    const bool stream = (reinterpret_cast<uintptr_t>(dst) & 3) == 0 && n >= 64;
    if (stream) {
        while ((reinterpret_cast<uintptr_t>(dst + k) & 31) != 0) {
            dst[k] = (int16_t)(src[0] | ((src[1] & 0x0f) << 8));
            dst[k + 1] = (int16_t)((src[1] >> 4) | (src[2] << 4));
            k += 2; src += 3;
        }
    }
    // the second 16-byte load of a step reads 4 bytes past the 24 it uses: stop 16 samples early
    for (; k + 32 <= n; k += 16, src += 24) {
        const __m128i lo = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
        const __m128i hi = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 12));
        const __m256i v = _mm256_shuffle_epi8(_mm256_inserti128_si256(_mm256_castsi128_si256(lo), hi, 1), shuf);
        const __m256i even = _mm256_and_si256(v, mask), odd = _mm256_srli_epi16(v, 4);
        const __m256i w = _mm256_blend_epi16(even, odd, 0xaa);
        if (stream) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + k), w);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + k), w);
    }
    if (stream) _mm_sfence();
    unpack12_scalar(src, dst + k, n - k);
}
#endif

inline void unpack12(const unsigned char* src, int16_t* dst, size_t n) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { unpack12_avx2(src, dst, n); return; }
#endif
    unpack12_scalar(src, dst, n);
}

// A fixed set of worker threads that run one "parallel for" at a time (the caller's thread takes part).
class WorkerPool {
public:
    explicit WorkerPool(int n_threads) {
        for (int i = 1; i < n_threads; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++epoch_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return (int)threads_.size() + 1; }
    // fn(part) for part = 0 .. n_parts - 1, spread over the pool; returns when all are done
    void run(int n_parts, const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn; n_parts_ = n_parts; next_.store(0); left_.store(n_parts); ++epoch_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return left_.load() == 0; });
        fn_ = nullptr;
    }

private:
    void work() {
        while (true) {
            const int part = next_.fetch_add(1);
            if (part >= n_parts_) break;
            (*fn_)(part);
            if (left_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> l(m_); done_.notify_all(); }
        }
    }
    void loop() {
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_parts_ = 0;
    std::atomic<int> next_{0}, left_{0};
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

// n samples, packed -> int16, in slices of whole sample pairs over the pool
inline void unpack12_parallel(WorkerPool& pool, const unsigned char* src, int16_t* dst, size_t n) {
    const size_t slice = (size_t)1 << 20;                        // samples per part (even)
    const int parts = (int)((n + slice - 1) / slice);
    pool.run(parts, [&](int part) {
        const size_t s0 = (size_t)part * slice, s1 = std::min(n, s0 + slice);
        unpack12(src + s0 / 2 * 3, dst + s0, s1 - s0);
    });
}

}  // namespace dct
