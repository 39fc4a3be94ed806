// ppb_pack.h -- run-length packed frame download (see ppb_pack.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace ppb {

struct PackPipe;

PackPipe *packpipe_create(int device);
void packpipe_destroy(PackPipe *p);   // waits for the jobs in flight
// Packs `n_frames` frames of `npx` RGBA pixels that the work already enqueued on `producer` leaves at `frames_dev`,
// and delivers them into `frames_host` asynchronously (job slot 0 or 1; blocks while the slot's previous job runs).
// `frames_dev` must stay untouched until packpipe_wait_slot(slot) returns.  Returns the number of kernels launched
// (0 for an empty job), -1 on error.
int packpipe_submit(PackPipe *p, int slot, const uint8_t *frames_dev, long long n_frames, int npx, uint8_t *frames_host,
                    cudaStream_t producer, std::string *err);
long long packpipe_submitted(PackPipe *p);
int packpipe_wait_finished(PackPipe *p, long long target, std::string *err);
int packpipe_wait_slot(PackPipe *p, int slot, std::string *err);
// out = {jobs that went packed, jobs that went raw, bytes copied packed (tokens + tables), bytes copied raw,
//        microseconds spent in the token copies, microseconds spent unpacking}
void packpipe_stats(PackPipe *p, long long out[6]);

}  // namespace ppb
