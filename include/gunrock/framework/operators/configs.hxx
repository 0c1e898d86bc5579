/**
 * @file configs.hxx
 * @brief Compile-time switches of the operators. Enumerator names and order are the reference's
 * (include/gunrock/framework/operators/configs.hxx:31-92) since user code names them in execute<...>().
 * What each load balancer means in this implementation is documented in advance/advance.hxx.
 */
#pragma once

namespace gunrock {
namespace operators {

enum load_balance_t { thread_mapped, warp_mapped, block_mapped, bucketing, merge_path, merge_path_v2, work_stealing };
enum advance_io_type_t { graph, vertices, edges, none };
enum advance_direction_t { forward, backward, optimized };
enum filter_algorithm_t { remove, predicated, compact, bypass };
enum uniquify_algorithm_t { unique, unique_copy };
enum parallel_for_each_t { vertex, edge, weight, element };

}  // namespace operators
}  // namespace gunrock
