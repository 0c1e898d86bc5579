/**
 * @file exchange.hxx
 * @brief operators::exchange::execute — the per-iteration frontier exchange of a 1-D partitioned (one process per
 * GPU) enactor: every vertex of the frontier an advance just produced travels to its owner together with its label,
 * the owner folds the label in with atomic::min and the vertices whose label improved form the owner's next frontier.
 *
 * The reference has no multi-GPU path: its operators throw on `context.size() != 1`
 * (include/gunrock/framework/operators/advance/advance.hxx:125-128) and enact() only uses get_context(0)
 * (framework/enactor.hxx:243-254); gcuda::multi_context_t(devices) / enable_peer_access()
 * (cuda/context.hxx:136-206) is the hook it leaves. Here a multi_context_t carries a partition descriptor
 * (gcuda::partition_t: rank, world, vertices per rank, three stream-ordered collectives bound by the host —
 * the C ABI binds NCCL over NVLink) and the SAME enactor contract runs partitioned:
 *
 *   loop():  advance::execute<lb>(G, E, op, context);          // unchanged lambda; G = owned rows, global column ids
 *            exchange::execute(G, E, labels, context);         // frontier -> owners, min-combine, dedupe
 *   is_converged(): the frontier is empty on EVERY rank (enactor_t::global_frontier_size)
 *
 * Conventions of a partitioned run (graph_properties_t::row_offset / global_vertices): frontiers hold LOCAL row ids;
 * operators receive GLOBAL vertex ids for source and neighbour, so label arrays are full-length and indexed
 * globally; after an exchange the owned entries of `labels` are exact, the others are this rank's best candidates
 * so far (a pruning bound only). BFS depths and SSSP distances are min-fixed points, so the owned slices equal the
 * single-GPU result bit for bit.
 *
 * Per call: bin the frontier by owner (shared-memory histogram, one global atomic per owner per CTA), all_gather of
 * the P counts, all_to_all_v of (id, label) records, absorb (atomic::min on owned labels, test-and-set of a per-level
 * bitmap, warp-aggregated append), all_reduce of the new frontier size: O(|frontier|) bytes over NVLink, two host
 * round trips. The direction-optimised BFS with bitmap exchange over peer memory (ess_dist_bfs) remains the fast
 * specialised driver; this operator is the general, operator-API form.
 */
#pragma once

#include <vector>

#include <gunrock/cuda/cuda.hxx>
#include <gunrock/error.hxx>
#include <gunrock/util/math.hxx>
#include <gunrock/util/type_limits.hxx>
#include <gunrock/b200/warp.cuh>
#include <gunrock/framework/operators/advance/kernels.cuh>

namespace gunrock {
namespace operators {
namespace exchange {

constexpr int max_ranks = 64;

namespace kernels {

using b200::counter_t;

/// counts[o] += number of valid frontier entries owned by rank o.
template <typename vertex_t>
__global__ void __launch_bounds__(256)
    count_owners_kernel(const vertex_t* __restrict__ frontier, std::size_t size, long long per, int world,
                        counter_t* counts) {
  __shared__ unsigned hist[max_ranks];
  if (threadIdx.x < max_ranks) hist[threadIdx.x] = 0;
  __syncthreads();
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < size;
       i += std::size_t(gridDim.x) * blockDim.x) {
    const vertex_t v = frontier[i];
    if (util::limits::is_valid(v)) atomicAdd(&hist[int((long long)v / per)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < unsigned(world) && hist[threadIdx.x]) atomicAdd(counts + threadIdx.x, counter_t(hist[threadIdx.x]));
}

/// records[cursor[o]++] = (v, labels[v]) for every valid frontier entry v owned by o. `cursor` starts at the send
/// offsets; a CTA reserves one range per owner per tile of 1024 entries (order inside a range is arbitrary).
template <typename vertex_t, typename label_t>
__global__ void __launch_bounds__(256)
    bin_records_kernel(const vertex_t* __restrict__ frontier, std::size_t size, const label_t* __restrict__ labels,
                       long long per, int world, counter_t* cursor, uint2* __restrict__ records) {
  static_assert(sizeof(label_t) == 4 && sizeof(vertex_t) == 4, "records are (32-bit id, 32-bit label)");
  __shared__ unsigned hist[max_ranks];
  __shared__ unsigned long long base[max_ranks];
  const std::size_t tile = 1024;
  for (std::size_t first = std::size_t(blockIdx.x) * tile; first < size; first += std::size_t(gridDim.x) * tile) {
    if (threadIdx.x < max_ranks) hist[threadIdx.x] = 0;
    __syncthreads();
    vertex_t v[4];
    int owner[4];
    unsigned slot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const std::size_t i = first + std::size_t(k) * 256 + threadIdx.x;
      v[k] = i < size ? frontier[i] : gunrock::numeric_limits<vertex_t>::invalid();
      owner[k] = -1;
      if (util::limits::is_valid(v[k])) {
        owner[k] = int((long long)v[k] / per);
        slot[k] = atomicAdd(&hist[owner[k]], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < unsigned(world) && hist[threadIdx.x])
      base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, counter_t(hist[threadIdx.x]));
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (owner[k] >= 0)
        records[base[owner[k]] + slot[k]] = make_uint2(unsigned(v[k]), __builtin_bit_cast(unsigned, labels[v[k]]));
    __syncthreads();
  }
}

/// Owner side. Records [own_begin, own_end) are this rank's own (their labels are already in place); every other
/// record folds its label in with atomic::min and counts as improved iff it lowered the label. An improved vertex is
/// appended (LOCAL id) the first time its bit in `seen` is set.
template <typename vertex_t, typename label_t>
__global__ void __launch_bounds__(256)
    absorb_records_kernel(const uint2* __restrict__ records, std::size_t count, std::size_t own_begin,
                          std::size_t own_end, label_t* __restrict__ labels, long long row_begin,
                          unsigned* __restrict__ seen, vertex_t* __restrict__ out, counter_t capacity,
                          counter_t* out_count) {
  const std::size_t stride = std::size_t(gridDim.x) * blockDim.x;
  const std::size_t rounds = (count + stride - 1) / stride;
  for (std::size_t r = 0; r < rounds; ++r) {  // warp-uniform trip count keeps the aggregated append converged
    const std::size_t i = r * stride + std::size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    bool keep = false;
    vertex_t local = 0;
    if (i < count) {
      const uint2 rec = records[i];
      const vertex_t v = vertex_t(rec.x);
      const label_t value = __builtin_bit_cast(label_t, rec.y);
      bool improved = true;
      if (i < own_begin || i >= own_end) improved = value < math::atomic::min(labels + v, value);
      if (improved) {
        local = vertex_t((long long)v - row_begin);
        const unsigned bit = 1u << (unsigned(local) & 31u);
        unsigned* word = seen + (unsigned(local) >> 5);
        keep = !(*word & bit) && !(atomicOr(word, bit) & bit);
      }
    }
    const counter_t at = b200::warp_append_slot(keep, out_count);
    if (keep && at < capacity) out[at] = local;
  }
}

}  // namespace kernels

/**
 * @brief Routes E's input frontier (GLOBAL ids, as the advance that just ran left it) to the owners; on return E's
 * input frontier holds the LOCAL ids of the owned vertices whose label improved and E->global_frontier_size the
 * frontier size summed over all ranks. With a single rank it only translates and de-duplicates.
 */
template <typename graph_t, typename enactor_type, typename label_t>
void execute(graph_t& G, enactor_type* E, label_t* labels, gcuda::multi_context_t& context, bool swap_buffers = true) {
  using vertex_t = typename graph_t::vertex_type;
  using kernels::counter_t;
  auto* part = context.partition.get();
  error::throw_if_exception(part == nullptr, "exchange::execute: the context carries no partition descriptor");
  const int world = part->world, rank = part->rank;
  error::throw_if_exception(world > max_ranks, "exchange::execute: too many ranks");
  auto* ctx = context.get_context(0);
  auto stream = ctx->stream();
  auto& scratch = ctx->scratch();
  auto* in = E->get_input_frontier();
  auto* out = E->get_output_frontier();
  const std::size_t nf = in->get_number_of_elements();
  const long long per = part->per, row_begin = per * rank;
  const std::size_t n_local = std::size_t(G.get_number_of_vertices());

  // ---- how many records go to each owner; every rank learns the whole P x P matrix ----
  memory::device_array_t<counter_t>& table = E->exchange_counts;
  table.resize(std::size_t(world) * world + 2 * world + 2);
  counter_t* mine = table.data();                 // [world]   my counts per owner, then reused as cursors
  counter_t* matrix = table.data() + world;       // [world*world]
  counter_t* total = matrix + std::size_t(world) * world;  // [2] new local / global frontier size
  cudaMemsetAsync(mine, 0, std::size_t(world) * sizeof(counter_t), stream);
  if (nf)
    kernels::count_owners_kernel<<<gcuda::persistent_grid(*ctx, (nf + 255) / 256, 4), 256, 0, stream>>>(
        in->data(), nf, per, world, mine);
  part->all_gather(mine, matrix, std::size_t(world) * sizeof(counter_t), stream);
  std::vector<counter_t> h(std::size_t(world) * world);
  cudaMemcpyAsync(h.data(), matrix, h.size() * sizeof(counter_t), cudaMemcpyDeviceToHost, stream);
  ctx->synchronize();
  std::vector<std::size_t> send_bytes(world), send_off(world), recv_bytes(world), recv_off(world);
  std::vector<counter_t> cursor(world);
  std::size_t n_send = 0, n_recv = 0;
  for (int p = 0; p < world; ++p) {
    send_off[p] = n_send * sizeof(uint2);
    cursor[p] = n_send;
    send_bytes[p] = std::size_t(h[std::size_t(rank) * world + p]) * sizeof(uint2);
    n_send += std::size_t(h[std::size_t(rank) * world + p]);
    recv_off[p] = n_recv * sizeof(uint2);
    recv_bytes[p] = std::size_t(h[std::size_t(p) * world + rank]) * sizeof(uint2);
    n_recv += std::size_t(h[std::size_t(p) * world + rank]);
  }
  // ---- bin (id, label) records by owner, ship them, fold them in ----
  E->exchange_send.reserve(n_send + 1, false);
  E->exchange_recv.reserve(n_recv + 1, false);
  uint2* send = E->exchange_send.data();
  uint2* recv = E->exchange_recv.data();
  cudaMemcpyAsync(mine, cursor.data(), std::size_t(world) * sizeof(counter_t), cudaMemcpyHostToDevice, stream);
  if (nf)
    kernels::bin_records_kernel<<<gcuda::persistent_grid(*ctx, (nf + 1023) / 1024, 4), 256, 0, stream>>>(
        in->data(), nf, labels, per, world, mine, send);
  part->all_to_all_v(send, send_bytes.data(), send_off.data(), recv, recv_bytes.data(), recv_off.data(), stream);
  auto& seen = E->unique_seen;
  if (seen.get_universe() != n_local) {
    seen.resize(n_local, stream);
    seen.fill(0, stream);
  }
  if (out->get_capacity() < n_local) out->reserve(n_local);
  scratch.zero(stream);
  if (n_recv)
    kernels::absorb_records_kernel<<<gcuda::persistent_grid(*ctx, (n_recv + 255) / 256, 4), 256, 0, stream>>>(
        recv, n_recv, recv_off[rank] / sizeof(uint2), (recv_off[rank] + recv_bytes[rank]) / sizeof(uint2), labels,
        row_begin, seen.data(), out->data(), counter_t(out->get_capacity()),
        scratch.d + gcuda::scratch_t::out_count);
  ctx->profiler().launches_total += 3;
  // new frontier size here and everywhere
  cudaMemcpyAsync(total, scratch.d + gcuda::scratch_t::out_count, sizeof(counter_t), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(total + 1, total, sizeof(counter_t), cudaMemcpyDeviceToDevice, stream);
  part->all_reduce_sum(reinterpret_cast<long long*>(total + 1), 1, stream);
  counter_t sizes[2] = {0, 0};
  cudaMemcpyAsync(sizes, total, 2 * sizeof(counter_t), cudaMemcpyDeviceToHost, stream);
  ctx->synchronize();
  scratch.clean = false;  // out_count was read without the publish kernel
  out->set_number_of_elements(std::size_t(sizes[0]));
  E->global_frontier_size = (long long)sizes[1];
  if (sizes[0])  // leave the per-level bitmap all clear again: O(|frontier|)
    advance::kernels::clear_emitted_kernel<<<gcuda::persistent_grid(*ctx, (std::size_t(sizes[0]) + 255) / 256, 4), 256,
                                             0, stream>>>(out->data(), std::size_t(sizes[0]), seen.data());
  if (swap_buffers) E->swap_frontier_buffers();
}

}  // namespace exchange
}  // namespace operators
}  // namespace gunrock
