// maxflow_wrap.cpp -- C wrapper around the reference's VENDORED Boykov-Kolmogorov max-flow (src/max_flow/graph.h,
// graph.cpp, maxflow.cpp, v3.04), compiled from the sources where they lie under /root/reference into
// oracle/_ref/libmaxflow.so (recipe: oracle/graph_cut.py build()).  TEST / INPUT-GENERATION INFRASTRUCTURE ONLY: it
// produces the graph-cut seam masks that are an INPUT of the hot path (BASELINE.json configs[4]); nothing of the product
// links it.  The call sequence is gcut::define_graph_full's (src/math/_graph_cut.cpp:344-405): one node per object
// pixel, then per node its horizontal edge, its vertical edge and its terminal weights, in node order -- the order
// decides how the search trees grow, so it is kept.
#include "graph.h"

extern "C" float mf_solve(int n, const int *h_conn, const float *h_cap, const int *v_conn, const float *v_cap,
                          const unsigned char *sink, const unsigned char *source, int n_edges, int *label)
{
    typedef Graph<float, float, float> G;
    G *g = new G(n, n_edges);
    for (int i = 0; i < n; ++i) g->add_node();
    for (int i = 0; i < n; ++i) {
        if (h_conn[i] >= 0) g->add_edge(i, h_conn[i], h_cap[i], h_cap[i]);
        if (v_conn[i] >= 0) g->add_edge(i, v_conn[i], v_cap[i], v_cap[i]);
        if (sink[i] || source[i]) g->add_tweights(i, (float)(sink[i] * 5000), (float)(source[i] * 5000));
    }
    const float flow = g->maxflow();
    for (int i = 0; i < n; ++i) label[i] = g->what_segment(i) == G::SOURCE;
    delete g;
    return flow;
}
