/* TEST INFRASTRUCTURE: the host-side worker / staging-job logic of libdbot_b200 (csrc/host_workers.cuh) on its
 * own, without CUDA: pageable buffers of many sizes are staged by the calling thread and helper threads that
 * claim pieces; every byte must arrive, no helper may be waited for that has not claimed anything, and workers
 * must run posted jobs exactly once.  Built and run by tests/test_host_workers.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>
static inline int cudaSetDevice(int) { return 0; } /* the header's only CUDA call (a shard worker's device) */
#include "../../dbot_ros_b200/csrc/host_workers.cuh"

int main()
{
    const int helpers = 3;
    std::vector<dbot_shard_worker*> w;
    for (int k = 0; k < helpers; ++k)
    {
        dbot_shard_worker* x = new dbot_shard_worker();
        x->spin = (k == 0) ? 40000 : (k == 1 ? 100 : 0); /* one spinning, one nearly, one always asleep */
        x->th = std::thread([x] { x->loop(-1); });
        w.push_back(x);
    }
    /* 1. posted jobs run exactly once, wait() returns after them */
    std::atomic<int> ran{0};
    for (int round = 0; round < 2000; ++round)
    {
        for (auto* x : w) x->post([&ran] { ran.fetch_add(1); });
        for (auto* x : w) x->wait();
        if (ran.load() != (round + 1) * helpers)
        {
            printf("FAIL: %d jobs ran after %d rounds\n", ran.load(), round + 1);
            return 1;
        }
    }
    /* 2. claim-based staging: every size, the caller never waits for a helper that holds no piece */
    dbot_stage_job* job = new dbot_stage_job();
    const size_t max_bytes = 640 * 480 * 4 + 77;
    std::vector<unsigned char> src(max_bytes), dst(max_bytes + 64);
    uint32_t seed = 12345u;
    long long pieces_by_helpers = 0, pieces_total = 0;
    for (int round = 0; round < 3000; ++round)
    {
        seed = seed * 1664525u + 1013904223u;
        size_t bytes = (round % 7 == 0) ? max_bytes : (seed >> 8) % max_bytes;
        if (round % 11 == 0) bytes = round % 5; /* almost nothing, and nothing */
        const int np = 1 + (int)((seed >> 3) % dbot_stage_job::MAX_PIECES);
        for (size_t i = 0; i < bytes; i += 997) src[i] = (unsigned char)(seed + i + round);
        memset(dst.data(), 0xee, bytes + 64);
        job->arm(dst.data(), src.data(), bytes, np);
        for (auto* x : w) x->post([job] { job->help(); });
        int mine = 0;
        for (int issued = 0; issued < job->np;)
        {
            if (job->claim_one()) ++mine;
            while (issued < job->np && job->is_done(issued)) ++issued;
        }
        pieces_total += np;
        pieces_by_helpers += np - mine;
        if (memcmp(dst.data(), src.data(), bytes) != 0)
        {
            printf("FAIL: round %d, %zu bytes in %d pieces differ\n", round, bytes, np);
            return 1;
        }
        for (int k = 0; k < 64; ++k)
            if (dst[bytes + k] != 0xee)
            {
                printf("FAIL: round %d wrote past the end\n", round);
                return 1;
            }
        if (job->cut(np) != bytes || job->cut(0) != 0)
        {
            printf("FAIL: cuts\n");
            return 1;
        }
    }
    for (auto* x : w)
    {
        x->wait(); /* late helpers find nothing left */
        x->stop();
        delete x;
    }
    delete job;
    printf("ok: %d jobs, %lld of %lld staging pieces taken by helpers\n", ran.load(), pieces_by_helpers, pieces_total);
    return 0;
}
