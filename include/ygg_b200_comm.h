/*
 * ygg_b200_comm.h — NCCL communicator of libygg_b200.so for the multi-GPU modes of ygg_b200.h
 * (one process per GPU, one node, NVLink/NVSwitch).
 *
 * The reference has no collective layer (its distributed GBT exchanges split records and evaluation
 * bitmaps as gRPC blobs between a manager and workers: learner/distributed_gradient_boosted_trees/
 * worker.proto:65-134, :199-203); this header is the B200 counterpart of that exchange, reduced to the
 * two collectives the level loop needs.  NCCL is resolved at run time with dlopen (libnccl.so.2 of the
 * process, e.g. the one PyTorch already loaded; override with YGG_B200_NCCL_LIB), so the library has no
 * link-time NCCL dependency and single-GPU hosts never touch it.
 *
 * Bootstrap: rank 0 calls ygg_comm_unique_id and ships the 128 bytes to the other ranks by any host-side
 * channel (torch.distributed broadcast, MPI, a file); every rank then calls ygg_comm_create.
 * ygg_comm_allreduce / ygg_comm_allgather have the callback signatures of ygg_b200.h, with the
 * communicator as `ctx`:
 *     ygg_gbt_set_row_shard(h, rank, world, n_global, init, ygg_comm_allreduce, comm);
 *     ygg_gbt_set_feature_shard(h, begin, end, rank, world, ygg_comm_allgather, comm);
 * They enqueue on the engine's stream and never synchronise with the host.
 */
#ifndef YGG_B200_COMM_H_
#define YGG_B200_COMM_H_

#include <stdint.h>

#include "ygg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define YGG_COMM_UNIQUE_ID_BYTES 128

typedef struct ygg_comm ygg_comm;

int ygg_comm_unique_id(uint8_t out[YGG_COMM_UNIQUE_ID_BYTES]);
int ygg_comm_create(ygg_comm** out, const uint8_t unique_id[YGG_COMM_UNIQUE_ID_BYTES], int32_t rank,
                    int32_t world, int32_t device);
int ygg_comm_destroy(ygg_comm* comm);
/* ygg_allreduce_fn / ygg_allgather_fn implementations (ctx = ygg_comm*). */
int ygg_comm_allreduce(void* ctx, void* buf, int64_t count, int32_t dtype, int32_t op, void* stream);
int ygg_comm_allgather(void* ctx, const void* send, void* recv, int64_t bytes, void* stream);
/* Peer-memory window (collective): `bytes` of zeroed device memory on every rank, mapped into every other rank's
 * address space with CUDA IPC (NVLink / NVSwitch peer access).  peers[r] receives THIS process's pointer to rank r's
 * window (peers[rank] = the local one).  The engine's kernels store into / spin on these windows directly
 * (ygg_gbt_set_best_split_window): the per-level best-split exchange then needs no collective call at all.
 * The windows live until ygg_comm_destroy. */
int ygg_comm_window_create(ygg_comm* comm, int64_t bytes, void** peers /* [world] */);
/* ygg_reducescatter_fn implementation (ncclReduceScatter, in place). */
int ygg_comm_reducescatter(void* ctx, void* buf, int64_t count_per_rank, int32_t dtype, int32_t op, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YGG_B200_COMM_H_ */
