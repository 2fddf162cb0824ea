// Collectives of the data-parallel path over NCCL (one rank per GPU, NVLink/NVSwitch):
//   utils.py:6-15    sync_networks  -> bmi_comm_bcast_f32 (root 0)
//   utils.py:43-48   sync_grads     -> bmi_comm_allreduce_sum_f32 (SUM, not averaged)
//   normalizer.py:60-64 _mpi_average -> bmi_comm_allreduce_sum_f32 + divide in bmi_norm_recompute
// NCCL calls only enqueue on the caller's stream, so they can sit inside the captured
// update graph between the backward pass and the Adam step.
#include <nccl.h>

#include "common.cuh"

struct bmi_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

#define BMI_NCCL_CHECK(expr)                                                                  \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      ::bmi::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r));  \
      return BMI_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)

using namespace bmi;

extern "C" int bmi_comm_unique_id(void* id128) {
  BMI_REQUIRE(id128, "bmi_comm_unique_id: null pointer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  BMI_NCCL_CHECK(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return BMI_OK;
}

extern "C" int bmi_comm_init(bmi_comm** out, int32_t rank, int32_t world, const void* id128) {
  BMI_REQUIRE(out && id128, "bmi_comm_init: null pointer");
  BMI_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bmi_comm_init: bad rank %d / world %d", rank, world);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  bmi_comm* c = new bmi_comm();
  c->rank = rank;
  c->world = world;
  ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank(rank %d / %d) -> %s", rank, world, ncclGetErrorString(r));
    delete c;
    return BMI_ERR_NCCL;
  }
  *out = c;
  return BMI_OK;
}

extern "C" int bmi_comm_destroy(bmi_comm* c) {
  if (!c) return BMI_OK;
  if (c->comm) ncclCommDestroy(c->comm);
  delete c;
  return BMI_OK;
}

extern "C" int bmi_comm_allreduce_sum_f32(bmi_comm* c, float* buf, int64_t n, bmi_stream_t stream) {
  BMI_REQUIRE(c && c->comm, "bmi_comm_allreduce_sum_f32: communicator not initialised");
  BMI_REQUIRE(n >= 0 && (buf || n == 0), "bmi_comm_allreduce_sum_f32: bad buffer");
  if (n == 0) return BMI_OK;
  BMI_NCCL_CHECK(ncclAllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, c->comm, as_stream(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return BMI_OK;
}

extern "C" int bmi_comm_bcast_f32(bmi_comm* c, float* buf, int64_t n, int32_t root, bmi_stream_t stream) {
  BMI_REQUIRE(c && c->comm, "bmi_comm_bcast_f32: communicator not initialised");
  BMI_REQUIRE(n >= 0 && (buf || n == 0) && root >= 0 && root < c->world, "bmi_comm_bcast_f32: bad arguments");
  if (n == 0) return BMI_OK;
  BMI_NCCL_CHECK(ncclBroadcast(buf, buf, (size_t)n, ncclFloat32, root, c->comm, as_stream(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return BMI_OK;
}
