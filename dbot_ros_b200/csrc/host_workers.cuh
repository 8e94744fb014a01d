/*
 * host_workers.cuh -- a host thread that runs posted jobs (shard enqueues in multi_device.cuh, staging copies of
 * pageable depth frames in dbot_b200.cu).  Part of libdbot_b200.so; included by dbot_b200.cu.
 */
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

/* Alive as long as the handle that owns it.  multi_device.cuh keeps one per shard beyond the first: it keeps its
 * GPU current and enqueues that shard's copies and kernels while the caller's thread does shard 0's (done from a
 * single thread, the last GPU of eight would start half a millisecond late; started per call, seven thread
 * creations cost a quarter of a millisecond).  A worker spins for the next job for `spin` polls after finishing
 * one (the calls of a frame come back to back), then sleeps on a condition variable. */
struct dbot_shard_worker
{
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    std::atomic<int> posted{0}, finished{0};
    bool quit = false;
    int spin = 40000;

    void loop(int device) /* device < 0: a plain host worker */
    {
        if (device >= 0) cudaSetDevice(device);
        int seen = 0;
        for (;;)
        {
            for (int k = 0; k < spin && posted.load(std::memory_order_acquire) == seen; ++k)
                __builtin_ia32_pause();
            if (posted.load(std::memory_order_acquire) == seen)
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || posted.load(std::memory_order_acquire) != seen; });
                if (quit) return;
            }
            std::function<void()> run;
            {
                /* a copy: post() may hand over the next job while this one is still running */
                std::lock_guard<std::mutex> lk(mu);
                run = job;
                seen = posted.load(std::memory_order_acquire);
            }
            run();
            finished.store(seen, std::memory_order_release);
        }
    }
    void post(std::function<void()> j)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(j);
            posted.fetch_add(1, std::memory_order_release);
        }
        cv.notify_one();
    }
    void wait()
    {
        const int want = posted.load(std::memory_order_acquire);
        while (finished.load(std::memory_order_acquire) != want) __builtin_ia32_pause();
    }
    void stop()
    {
        if (!th.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_one();
        th.join();
    }
};


/* A pageable buffer copied into a page-locked one in pieces that ANY thread may claim: the caller arms the job,
 * wakes the helpers and starts claiming pieces itself.  Helpers that wake promptly take their share; one that
 * wakes late (a sleeping thread can take a quarter of a millisecond to get a core on a busy host) finds nothing
 * left and costs nothing -- the caller never waits for a helper that has not claimed anything.  The job lives in
 * the handle, not on the stack: a late helper may look at it after the call has returned (it then sees every piece
 * taken, or already helps with the next frame). */
struct dbot_stage_job
{
    static constexpr int MAX_PIECES = 16;
    std::atomic<int> next{MAX_PIECES};
    std::atomic<int> done[MAX_PIECES];
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0;
    int np = 0;

    size_t cut(int i) const { return i >= np ? bytes : (bytes / 64 * (size_t)i / (size_t)np) * 64; }
    void arm(void* d, const void* s, size_t n, int pieces)
    {
        dst = (char*)d;
        src = (const char*)s;
        bytes = n;
        np = pieces;
        for (int i = 0; i < pieces; ++i) done[i].store(0, std::memory_order_relaxed);
        next.store(0, std::memory_order_release);
    }
    bool claim_one()
    {
        const int i = next.fetch_add(1, std::memory_order_acq_rel);
        if (i < 0 || i >= np) return false;
        memcpy(dst + cut(i), src + cut(i), cut(i + 1) - cut(i));
        done[i].store(1, std::memory_order_release);
        return true;
    }
    void help()
    {
        while (claim_one()) {}
    }
    bool is_done(int i) const { return done[i].load(std::memory_order_acquire) != 0; }
};
