// A small pool of host threads owned by a batch handle: moves rows from the pinned bounce buffers
// the GPU copies into to the caller's ordinary (pageable) arrays while the next chunk is computed
// and copied. A std::vector<StateVector> is pageable memory: a direct cudaMemcpy into it is
// staged by the driver through one thread (~10 GB/s); a handful of threads keep up with the
// PCIe link instead. C++11, no exceptions cross the C ABI (thread creation failures degrade to
// running the work on the calling thread).
#ifndef PERTURB_B200_HOST_WORKERS_H
#define PERTURB_B200_HOST_WORKERS_H

#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pb {

class HostWorkers {
public:
    explicit HostWorkers(unsigned threads);
    ~HostWorkers();
    HostWorkers(const HostWorkers &) = delete;
    HostWorkers &operator=(const HostWorkers &) = delete;

    // Calls fn(begin, end) over a partition of [0, n) on the pool and the calling thread;
    // returns when every part is done. Not re-entrant (one caller at a time, like the handle).
    void parallel_for(size_t n, const std::function<void(size_t, size_t)> &fn);

    unsigned size() const { return static_cast<unsigned>(threads_.size()) + 1; }

private:
    void worker();

    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(size_t, size_t)> *fn_;
    size_t n_, grain_, next_;
    unsigned long generation_;
    unsigned active_;
    bool stop_;
};

// Default pool size of ONE handle: half the CPUs this process may run on, shared between as many
// handles as the box has GPUs (one process per GPU, or one process with a handle per GPU: either
// way that many pools scatter at once, and on 8 GPUs each handle's share of the link is an
// eighth); at least 2, at most 16 (the scatter is memory bound: more threads only fight over the
// memory controllers).
unsigned default_host_workers(int visible_devices);

}  // namespace pb

#endif
