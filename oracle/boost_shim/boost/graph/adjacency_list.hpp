// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for <boost/graph/adjacency_list.hpp> (Boost is not installed in this
// image), just enough for the reference's src/CellGraph.cpp to compile UNMODIFIED into oracle/_ref, so that the edge
// loop of CellGraph::CellGraph (src/CellGraph.cpp:60-112) can produce golden vectors.  Written from the documented
// contract of the Boost Graph Library, not from its sources:
//   * add_vertex / add_edge append; vertices(g), edges(g), out_edges(v, g) iterate in insertion order (listS / vecS);
//   * in an undirected graph an edge {u, v} is an out-edge of both u and v; edge(u, v, g) reports whether one exists;
//   * with setS as the out-edge container parallel edges are not created (add_edge returns the existing one, false);
//   * descriptors: vertex = node pointer (listS) or index (vecS); g[v], g[e] give the bundled properties.
#pragma once
#include <cstddef>
#include <deque>
#include <list>
#include <tuple>
#include <utility>

namespace boost {

struct listS {};
struct vecS {};
struct setS {};
struct undirectedS {};
struct no_property {};

namespace shim {

template <class VertexS> struct VertexStore;

template <class G> struct EdgeNodeT {
    typename G::vertex_descriptor s, t;
    typename G::edge_property_type property;
};

}  // namespace shim

template <class OutEdgeS, class VertexS, class DirS, class VP = no_property, class EP = no_property>
class adjacency_list {
public:
    typedef adjacency_list Self;
    typedef VP vertex_property_type;
    typedef EP edge_property_type;
    static const bool kVec = sizeof(VertexS) == sizeof(vecS) && false;   // resolved through the trait below
    struct EdgeNode;
    struct VertexNode;
    // vertex descriptor: pointer for listS, index for vecS
    template <class S, class Dummy = void> struct Desc {
        typedef VertexNode* type;
    };
    template <class Dummy> struct Desc<vecS, Dummy> {
        typedef std::size_t type;
    };
    typedef typename Desc<VertexS>::type vertex_descriptor;
    struct edge_descriptor {
        vertex_descriptor s, t;
        EdgeNode* e;
        edge_descriptor() : s(), t(), e(nullptr) {}
        edge_descriptor(vertex_descriptor s_, vertex_descriptor t_, EdgeNode* e_) : s(s_), t(t_), e(e_) {}
        bool operator==(const edge_descriptor& o) const { return e == o.e; }
        bool operator!=(const edge_descriptor& o) const { return e != o.e; }
    };
    struct EdgeNode {
        vertex_descriptor s, t;
        EP property;
    };
    struct OutEdge {
        vertex_descriptor target;
        EdgeNode* e;
    };
    struct VertexNode {
        VP property;
        std::list<OutEdge> out;
        typename std::list<VertexNode>::iterator self;   // listS only
    };

    adjacency_list() {}
    explicit adjacency_list(std::size_t n)
    {
        for (std::size_t i = 0; i < n; i++) addVertex(VP());
    }

    static vertex_descriptor null_vertex() { return nullVertex(static_cast<VertexS*>(nullptr)); }

    VP& operator[](vertex_descriptor v) { return node(v).property; }
    const VP& operator[](vertex_descriptor v) const { return node(v).property; }
    EP& operator[](const edge_descriptor& e) { return e.e->property; }
    const EP& operator[](const edge_descriptor& e) const { return e.e->property; }

    // ---- iterators ------------------------------------------------------------------------------
    class vertex_iterator {
    public:
        vertex_iterator() : g(nullptr), i(0) {}
        vertex_descriptor operator*() const { return g->descriptorAt(it, i); }
        vertex_iterator& operator++()
        {
            ++i;
            if (it != g->vertexList.end()) ++it;
            return *this;
        }
        bool operator!=(const vertex_iterator& o) const { return i != o.i; }
        bool operator==(const vertex_iterator& o) const { return i == o.i; }
        const Self* g;
        typename std::list<VertexNode>::const_iterator it;
        std::size_t i;
    };
    class edge_iterator {
    public:
        edge_descriptor operator*() const { return edge_descriptor(it->s, it->t, const_cast<EdgeNode*>(&*it)); }
        edge_iterator& operator++()
        {
            ++it;
            return *this;
        }
        bool operator!=(const edge_iterator& o) const { return it != o.it; }
        bool operator==(const edge_iterator& o) const { return it == o.it; }
        typename std::list<EdgeNode>::const_iterator it;
    };
    class out_edge_iterator {
    public:
        edge_descriptor operator*() const { return edge_descriptor(v, it->target, it->e); }
        out_edge_iterator& operator++()
        {
            ++it;
            return *this;
        }
        bool operator!=(const out_edge_iterator& o) const { return it != o.it; }
        bool operator==(const out_edge_iterator& o) const { return it == o.it; }
        vertex_descriptor v;
        typename std::list<OutEdge>::const_iterator it;
    };

    // ---- storage --------------------------------------------------------------------------------
    std::list<VertexNode> vertexList;          // listS
    std::deque<VertexNode> vertexVec;          // vecS
    std::list<EdgeNode> edgeList;

    VertexNode& node(vertex_descriptor v) { return nodeOf(v, static_cast<VertexS*>(nullptr)); }
    const VertexNode& node(vertex_descriptor v) const { return const_cast<Self*>(this)->nodeOf(v, static_cast<VertexS*>(nullptr)); }
    std::size_t vertexCount() const { return isVec(static_cast<VertexS*>(nullptr)) ? vertexVec.size() : vertexList.size(); }

    vertex_descriptor addVertex(const VP& p) { return addVertexImpl(p, static_cast<VertexS*>(nullptr)); }
    void removeVertex(vertex_descriptor v) { removeVertexImpl(v, static_cast<VertexS*>(nullptr)); }
    vertex_descriptor descriptorAt(typename std::list<VertexNode>::const_iterator it, std::size_t i) const
    {
        return descriptorAtImpl(it, i, static_cast<VertexS*>(nullptr));
    }

private:
    static bool isVec(vecS*) { return true; }
    template <class S> static bool isVec(S*) { return false; }
    static std::size_t nullVertex(vecS*) { return std::size_t(-1); }
    template <class S> static VertexNode* nullVertex(S*) { return nullptr; }
    VertexNode& nodeOf(std::size_t v, vecS*) { return vertexVec[v]; }
    template <class S> VertexNode& nodeOf(VertexNode* v, S*) { return *v; }
    std::size_t addVertexImpl(const VP& p, vecS*)
    {
        vertexVec.emplace_back();
        vertexVec.back().property = p;
        return vertexVec.size() - 1;
    }
    template <class S> VertexNode* addVertexImpl(const VP& p, S*)
    {
        vertexList.emplace_back();
        vertexList.back().property = p;
        vertexList.back().self = --vertexList.end();
        return &vertexList.back();
    }
    void removeVertexImpl(std::size_t, vecS*) {}
    template <class S> void removeVertexImpl(VertexNode* v, S*) { vertexList.erase(v->self); }
    std::size_t descriptorAtImpl(typename std::list<VertexNode>::const_iterator, std::size_t i, vecS*) const { return i; }
    template <class S> VertexNode* descriptorAtImpl(typename std::list<VertexNode>::const_iterator it, std::size_t, S*) const
    {
        return const_cast<VertexNode*>(&*it);
    }
};

#define EM2_SHIM_G adjacency_list<O, V, D, VP, EP>
#define EM2_SHIM_T template <class O, class V, class D, class VP, class EP>

EM2_SHIM_T typename EM2_SHIM_G::vertex_descriptor add_vertex(const VP& p, EM2_SHIM_G& g) { return g.addVertex(p); }
EM2_SHIM_T typename EM2_SHIM_G::vertex_descriptor add_vertex(EM2_SHIM_G& g) { return g.addVertex(VP()); }

EM2_SHIM_T std::pair<typename EM2_SHIM_G::edge_descriptor, bool> edge(typename EM2_SHIM_G::vertex_descriptor u,
                                                                       typename EM2_SHIM_G::vertex_descriptor v, const EM2_SHIM_G& g)
{
    typedef typename EM2_SHIM_G::edge_descriptor E;
    for (const auto& o : g.node(u).out)
        if (o.target == v) return std::make_pair(E(u, v, o.e), true);
    return std::make_pair(E(), false);
}

namespace shim {
inline bool uniqueOutEdges(setS*) { return true; }
template <class S> bool uniqueOutEdges(S*) { return false; }
}  // namespace shim

EM2_SHIM_T std::pair<typename EM2_SHIM_G::edge_descriptor, bool> add_edge(typename EM2_SHIM_G::vertex_descriptor u,
                                                                           typename EM2_SHIM_G::vertex_descriptor v, const EP& p,
                                                                           EM2_SHIM_G& g)
{
    typedef typename EM2_SHIM_G::edge_descriptor E;
    if (shim::uniqueOutEdges(static_cast<O*>(nullptr))) {
        const auto found = edge(u, v, g);
        if (found.second) return std::make_pair(found.first, false);
    }
    g.edgeList.emplace_back();
    typename EM2_SHIM_G::EdgeNode* e = &g.edgeList.back();
    e->s = u;
    e->t = v;
    e->property = p;
    g.node(u).out.push_back({v, e});
    if (!(u == v)) g.node(v).out.push_back({u, e});
    return std::make_pair(E(u, v, e), true);
}
EM2_SHIM_T std::pair<typename EM2_SHIM_G::edge_descriptor, bool> add_edge(typename EM2_SHIM_G::vertex_descriptor u,
                                                                           typename EM2_SHIM_G::vertex_descriptor v, EM2_SHIM_G& g)
{
    return add_edge(u, v, EP(), g);
}

EM2_SHIM_T std::pair<typename EM2_SHIM_G::vertex_iterator, typename EM2_SHIM_G::vertex_iterator> vertices(const EM2_SHIM_G& g)
{
    typename EM2_SHIM_G::vertex_iterator b, e;
    b.g = e.g = &g;
    b.it = g.vertexList.begin();
    e.it = g.vertexList.end();
    b.i = 0;
    e.i = g.vertexCount();
    return std::make_pair(b, e);
}
EM2_SHIM_T std::pair<typename EM2_SHIM_G::edge_iterator, typename EM2_SHIM_G::edge_iterator> edges(const EM2_SHIM_G& g)
{
    typename EM2_SHIM_G::edge_iterator b, e;
    b.it = g.edgeList.begin();
    e.it = g.edgeList.end();
    return std::make_pair(b, e);
}
EM2_SHIM_T std::pair<typename EM2_SHIM_G::out_edge_iterator, typename EM2_SHIM_G::out_edge_iterator> out_edges(
    typename EM2_SHIM_G::vertex_descriptor v, const EM2_SHIM_G& g)
{
    typename EM2_SHIM_G::out_edge_iterator b, e;
    b.v = e.v = v;
    b.it = g.node(v).out.begin();
    e.it = g.node(v).out.end();
    return std::make_pair(b, e);
}
EM2_SHIM_T std::size_t out_degree(typename EM2_SHIM_G::vertex_descriptor v, const EM2_SHIM_G& g) { return g.node(v).out.size(); }
EM2_SHIM_T std::size_t num_vertices(const EM2_SHIM_G& g) { return g.vertexCount(); }
EM2_SHIM_T std::size_t num_edges(const EM2_SHIM_G& g) { return g.edgeList.size(); }
EM2_SHIM_T typename EM2_SHIM_G::vertex_descriptor source(const typename EM2_SHIM_G::edge_descriptor& e, const EM2_SHIM_G&) { return e.s; }
EM2_SHIM_T typename EM2_SHIM_G::vertex_descriptor target(const typename EM2_SHIM_G::edge_descriptor& e, const EM2_SHIM_G&) { return e.t; }
// the caller has removed the vertex's edges already (BGL's precondition; CellGraph only removes isolated vertices)
EM2_SHIM_T void remove_vertex(typename EM2_SHIM_G::vertex_descriptor v, EM2_SHIM_G& g) { g.removeVertex(v); }

// property map of a bundled member: only handed to write_graphviz (a stub here)
template <class T, class C> struct shim_member_map {
    T C::*member;
};
EM2_SHIM_T struct graph_traits_shim {};
template <class T, class C, class O, class V, class D, class VP, class EP>
shim_member_map<T, C> get(T C::*m, const EM2_SHIM_G&)
{
    return shim_member_map<T, C>{m};
}

// boost::tie (found by argument-dependent lookup at src/CellGraph.cpp:106)
template <class A, class B> std::tuple<A&, B&> tie(A& a, B& b) { return std::tuple<A&, B&>(a, b); }

template <class G> struct graph_traits {
    typedef typename G::vertex_descriptor vertex_descriptor;
    typedef typename G::edge_descriptor edge_descriptor;
    typedef typename G::vertex_iterator vertex_iterator;
    typedef typename G::edge_iterator edge_iterator;
    typedef typename G::out_edge_iterator out_edge_iterator;
};

#undef EM2_SHIM_G
#undef EM2_SHIM_T

}  // namespace boost
