// See host_workers.h.
#include "host_workers.h"

#include <algorithm>
#ifdef __linux__
#include <sched.h>
#endif

namespace pb {

unsigned default_host_workers(int visible_devices) {
    unsigned hw = std::thread::hardware_concurrency();
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) {
        hw = static_cast<unsigned>(CPU_COUNT(&set));  // taskset / cgroup limits count
    }
#endif
    if (hw == 0) hw = 2;
    const unsigned pools = static_cast<unsigned>(std::max(1, visible_devices));
    return std::min(16u, std::max(2u, hw / (2 * pools)));
}

HostWorkers::HostWorkers(unsigned threads)
    : fn_(nullptr), n_(0), grain_(1), next_(0), generation_(0), active_(0), stop_(false) {
    for (unsigned i = 1; i < threads; ++i) {
        try {
            threads_.emplace_back(&HostWorkers::worker, this);
        } catch (...) {
            break;  // fewer helpers (possibly none): parallel_for still completes
        }
    }
}

HostWorkers::~HostWorkers() {
    {
        std::lock_guard<std::mutex> lock(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (std::thread &t : threads_) t.join();
}

void HostWorkers::worker() {
    unsigned long seen = 0;
    std::unique_lock<std::mutex> lock(mu_);
    for (;;) {
        cv_work_.wait(lock, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        while (next_ < n_) {
            const size_t begin = next_;
            const size_t end = std::min(n_, begin + grain_);
            next_ = end;
            ++active_;
            lock.unlock();
            (*fn_)(begin, end);
            lock.lock();
            --active_;
        }
        if (active_ == 0) cv_done_.notify_all();
    }
}

void HostWorkers::parallel_for(size_t n, const std::function<void(size_t, size_t)> &fn) {
    if (n == 0) return;
    if (threads_.empty()) {
        fn(0, n);
        return;
    }
    std::unique_lock<std::mutex> lock(mu_);
    fn_ = &fn;
    n_ = n;
    // a few parts per thread so that an unlucky thread does not hold everyone up
    grain_ = std::max<size_t>(1, n / (4 * (threads_.size() + 1)));
    next_ = 0;
    ++generation_;
    cv_work_.notify_all();
    while (next_ < n_) {  // the calling thread works too
        const size_t begin = next_;
        const size_t end = std::min(n_, begin + grain_);
        next_ = end;
        ++active_;
        lock.unlock();
        fn(begin, end);
        lock.lock();
        --active_;
    }
    cv_done_.wait(lock, [&] { return active_ == 0; });
    fn_ = nullptr;
    n_ = 0;
}

}  // namespace pb
