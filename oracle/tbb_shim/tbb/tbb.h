// TEST INFRASTRUCTURE ONLY (oracle/): a std::thread stand-in for the six Intel TBB symbols
// that the *reference's generated C++* uses, so that code can be built in an image that has no
// libtbb (no network).  It supplies scheduling only -- no query arithmetic lives here.
//
// Symbols and where the reference emits them (file:line in /root/reference/src/sdqlpy/lib):
//   tbb::task_scheduler_init        sdql_compiler.py:485
//   tbb::blocked_range<size_t>      sdql_ir_cpp_generator_par.py:206-210
//   tbb::parallel_for               sdql_ir_cpp_generator_par.py:206, 311, 349, 420
//   tbb::parallel_reduce            sdql_ir_cpp_generator_par.py:270-288
//   tbb::enumerable_thread_specific sdql_ir_cpp_generator_par.py:310, 348, 419
//   tbb::concurrent_vector          sdql_ir_cpp_generator_par.py:203-204
//
// Anything timed through this header must be labelled "TBB-shim", never "TBB".
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tbb {

namespace shim {
inline int& nthreads() {
    static int n = 1;
    return n;
}
// Run fn(chunk_begin, chunk_end, chunk_index) over [b, e) on nthreads() threads, chunks handed
// out dynamically (a coarse imitation of TBB's work stealing).
template <class F>
inline size_t run_chunks(size_t b, size_t e, size_t nchunks, F&& fn) {
    if (e <= b) return 0;
    size_t n = e - b;
    nchunks = std::max<size_t>(1, std::min(nchunks, n));
    size_t step = (n + nchunks - 1) / nchunks;
    nchunks = (n + step - 1) / step;
    int nt = std::min<size_t>(nthreads(), nchunks);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            size_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            size_t cb = b + c * step, ce = std::min(e, cb + step);
            fn(cb, ce, c);
        }
    };
    std::vector<std::thread> ths;
    for (int t = 1; t < nt; ++t) ths.emplace_back(worker);
    worker();
    for (auto& t : ths) t.join();
    return nchunks;
}
}  // namespace shim

struct task_scheduler_init {
    // thread count: the generator bakes N in at compile time (sdql_compiler.py:485); SDQL_REF_THREADS overrides it
    // at load time so one multi-threaded build can be timed on whatever core count the box has.
    explicit task_scheduler_init(int n) {
        const char* e = std::getenv("SDQL_REF_THREADS");
        if (e && std::atoi(e) > 0) n = std::atoi(e);
        shim::nthreads() = n > 0 ? n : 1;
    }
};

template <class T>
class blocked_range {
    T b_, e_;
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
};

template <class R, class F>
inline void parallel_for(const R& r, const F& f) {
    shim::run_chunks(r.begin(), r.end(), (size_t)shim::nthreads() * 8,
                     [&](size_t cb, size_t ce, size_t) { f(R(cb, ce)); });
}

template <class R, class T, class Body, class Join>
inline T parallel_reduce(const R& r, const T& identity, const Body& body, const Join& join) {
    size_t want = (size_t)shim::nthreads() * 8;
    std::vector<T> parts(want, identity);
    std::vector<char> used(want, 0);
    shim::run_chunks(r.begin(), r.end(), want, [&](size_t cb, size_t ce, size_t c) {
        parts[c] = body(R(cb, ce), identity);
        used[c] = 1;
    });
    T acc = identity;
    for (size_t i = 0; i < want; ++i)
        if (used[i]) acc = join(acc, parts[i]);
    return acc;
}

template <class T>
class enumerable_thread_specific {
    std::deque<T> items_;
    std::unordered_map<std::thread::id, T*> map_;
    std::mutex mu_;
public:
    T& local() {
        std::lock_guard<std::mutex> g(mu_);
        auto id = std::this_thread::get_id();
        auto it = map_.find(id);
        if (it != map_.end()) return *it->second;
        items_.emplace_back();
        map_[id] = &items_.back();
        return items_.back();
    }
    typename std::deque<T>::iterator begin() { return items_.begin(); }
    typename std::deque<T>::iterator end() { return items_.end(); }
};

template <class T>
class concurrent_vector {
    std::vector<T> v_;
    std::atomic_flag lock_ = ATOMIC_FLAG_INIT;
public:
    concurrent_vector() = default;
    concurrent_vector(const concurrent_vector& o) : v_(o.v_) {}
    concurrent_vector& operator=(const concurrent_vector& o) { v_ = o.v_; return *this; }
    template <class... A>
    void emplace_back(A&&... a) {
        while (lock_.test_and_set(std::memory_order_acquire)) {}
        v_.emplace_back(std::forward<A>(a)...);
        lock_.clear(std::memory_order_release);
    }
    size_t size() const { return v_.size(); }
    auto begin() const { return v_.begin(); }
    auto end() const { return v_.end(); }
};

}  // namespace tbb
