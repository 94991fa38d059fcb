// host_rng_service.h -- bulk TranscriptRng draws of concurrent provers, batched into SIMD lanes.
//
// A byte-exact proof needs 2n sequential 64-byte draws from its own merlin TranscriptRng (s_L, s_R of
// bulletproofs' Prover::prove, reached from /root/reference/src/bin/prover.rs:93): one Keccak-f each, ~50 ms of one
// host core at n = 63 180.  The streams of different proofs are independent, so provers running in the same process
// hand their bulk draws to this service: the first thread that finds no leader becomes one and advances up to W
// streams (its own and those queued by other threads) in the lanes of one AVX2 / AVX-512 register file
// (host_keccak_lanes.cpp), slice by slice, adopting newly queued streams into free lanes between slices.  The other
// threads sleep (their cores stay free for launching kernels).  When the leader's own stream is finished it puts the
// unfinished ones back and another waiter takes over.  The bytes are those of host_merlin.h's scalar TranscriptRng.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>

#include "host_merlin.h"

extern "C" void bpg_rng_lanes_avx2(uint64_t *st, uint8_t *const *out, const size_t *stride, size_t steps);
extern "C" void bpg_rng_lanes_avx512(uint64_t *st, uint8_t *const *out, const size_t *stride, size_t steps);

namespace bpgh {

class RngService {
    typedef void (*lanes_fn)(uint64_t *, uint8_t *const *, const size_t *, size_t);
    struct Job { uint64_t st[25]; uint8_t *out; size_t remaining; bool adopted, done; };
    enum { MAXW = 8, SLICE = 512 };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job *> queue;
    int leaders = 0;
    int W = 1;
    lanes_fn fn = nullptr;

    RngService() {
        int want = 8;
        if (const char *e = getenv("BPG_RNG_LANES")) want = atoi(e);
        __builtin_cpu_init();
        if (want >= 8 && __builtin_cpu_supports("avx512f")) { W = 8; fn = bpg_rng_lanes_avx512; }
        else if (want >= 4 && __builtin_cpu_supports("avx2")) { W = 4; fn = bpg_rng_lanes_avx2; }
    }

    // called with the lock held; returns with the lock held
    void lead(std::unique_lock<std::mutex> &lk, Job *own) {
        leaders++;
        Job *lane[MAXW] = {nullptr};
        alignas(64) uint64_t soa[25 * MAXW];
        memset(soa, 0, sizeof soa);
        uint8_t dummy[64];
        // the leader's own stream first
        for (auto it = queue.begin(); it != queue.end(); ++it)
            if (*it == own) { queue.erase(it); break; }
        own->adopted = true;
        lane[0] = own;
        for (int k = 0; k < 25; k++) soa[k * W] = own->st[k];
        while (true) {
            for (int l = 0; l < W && !queue.empty(); l++) {
                if (lane[l]) continue;
                Job *j = queue.front();
                queue.pop_front();
                j->adopted = true;
                lane[l] = j;
                for (int k = 0; k < 25; k++) soa[k * W + l] = j->st[k];
            }
            size_t steps = SLICE;
            int active = 0;
            for (int l = 0; l < W; l++)
                if (lane[l]) { active++; if (lane[l]->remaining < steps) steps = lane[l]->remaining; }
            if (!active) break;
            uint8_t *outs[MAXW];
            size_t stride[MAXW];
            for (int l = 0; l < W; l++) { outs[l] = lane[l] ? lane[l]->out : dummy; stride[l] = lane[l] ? 64 : 0; }
            lk.unlock();
            fn(soa, outs, stride, steps);
            lk.lock();
            bool finished = false;
            for (int l = 0; l < W; l++) {
                Job *j = lane[l];
                if (!j) continue;
                j->out += 64 * steps;
                j->remaining -= steps;
                if (j->remaining == 0) {
                    for (int k = 0; k < 25; k++) j->st[k] = soa[k * W + l];
                    j->done = true;
                    lane[l] = nullptr;
                    finished = true;
                }
            }
            if (finished) cv.notify_all();
            if (own->done) { // hand the unfinished streams back (in lane order, ahead of newer arrivals)
                for (int l = W - 1; l >= 0; l--) {
                    Job *j = lane[l];
                    if (!j) continue;
                    for (int k = 0; k < 25; k++) j->st[k] = soa[k * W + l];
                    j->adopted = false;
                    queue.push_front(j);
                    lane[l] = nullptr;
                }
                break;
            }
        }
        explicit_bzero(soa, sizeof soa);
        leaders--;
        cv.notify_all();
    }

public:
    static RngService &get() { static RngService s; return s; }
    int lanes() const { return W; }

    // `count` consecutive rng.fill_bytes(out + 64 i, 64)
    void draw64(TranscriptRng &rng, uint8_t *out, size_t count) {
        Strobe &s = rng.s;
        // reach the steady state (position 64, no operation begun) with scalar draws; one is enough after any 64-byte draw
        while (count && (W == 1 || s.pos != 64 || s.pos_begin != 0)) { rng.fill_bytes(out, 64); out += 64; count--; }
        if (!count) return;
        Job job;
        memcpy(job.st, s.st, sizeof job.st);
        job.out = out; job.remaining = count; job.adopted = false; job.done = false;
        std::unique_lock<std::mutex> lk(mu);
        queue.push_back(&job);
        int waits = 0;
        while (!job.done) {
            // lead if nobody does, or if the current leaders have had no free lane for this stream for a whole slice
            // (the timed wait is only a safety net for the hand-over -- finished streams and retiring leaders notify; with
            // hundreds of waiting prover threads on a 32-core host a 0.5 ms poll cost about two cores)
            if (!job.adopted && (leaders == 0 || waits >= 2)) { lead(lk, &job); continue; }
            cv.wait_for(lk, std::chrono::microseconds(1000));
            waits++;
        }
        memcpy(s.st, job.st, sizeof job.st); // position and flags are those of the steady state again
        explicit_bzero(job.st, sizeof job.st); // the stream state is a prover secret
    }
};

} // namespace bpgh
