// TEST INFRASTRUCTURE ONLY -- stand-in for <boost/graph/iteration_macros.hpp> (see adjacency_list.hpp next to it).
#pragma once
#define BGL_FORALL_VERTICES(VNAME, GNAME, GraphType)                                                          \
    for (auto VNAME##_range = vertices(GNAME); VNAME##_range.first != VNAME##_range.second; VNAME##_range.first = VNAME##_range.second) \
        for (typename std::remove_reference<decltype(GNAME)>::type::vertex_descriptor VNAME;                   \
             VNAME##_range.first != VNAME##_range.second ? (VNAME = *VNAME##_range.first, true) : false; ++VNAME##_range.first)
#define BGL_FORALL_EDGES(ENAME, GNAME, GraphType)                                                             \
    for (auto ENAME##_range = edges(GNAME); ENAME##_range.first != ENAME##_range.second; ENAME##_range.first = ENAME##_range.second) \
        for (typename std::remove_reference<decltype(GNAME)>::type::edge_descriptor ENAME;                     \
             ENAME##_range.first != ENAME##_range.second ? (ENAME = *ENAME##_range.first, true) : false; ++ENAME##_range.first)
#define BGL_FORALL_OUTEDGES(UNAME, ENAME, GNAME, GraphType)                                                   \
    for (auto ENAME##_range = out_edges(UNAME, GNAME); ENAME##_range.first != ENAME##_range.second; ENAME##_range.first = ENAME##_range.second) \
        for (typename std::remove_reference<decltype(GNAME)>::type::edge_descriptor ENAME;                     \
             ENAME##_range.first != ENAME##_range.second ? (ENAME = *ENAME##_range.first, true) : false; ++ENAME##_range.first)
#include <type_traits>
