/** @file configs.hxx  Frontier enums; names and order as in the reference
 *  (include/gunrock/framework/frontier/configs.hxx:20-34) because user code spells them. */
#pragma once
namespace gunrock {
namespace frontier {
enum frontier_view_t { vector, bitmap, boolmap };
enum frontier_kind_t { vertex_frontier, edge_frontier, vertex_edge_frontier };
}  // namespace frontier
}  // namespace gunrock
