// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// C-ABI shim over the UNMODIFIED reference search core, compiled from the
// headers where they lie (-I$REF/poreover/decoding, see oracle/Makefile).
// Output goes to oracle/_ref/libporeover_ref.so (git-ignored).
//
// The reference's stock functions return only the decoded string
// (BeamSearch.h:400-458).  To also obtain the ranking score of the returned
// node WITHOUT restating any search loop, the reference's own function
// templates (beam_search_ BeamSearch.h:18, beam_search_2d_by_row :110/:175,
// beam_search_2d_by_row_col :262) are instantiated with a TBeam that derives
// from the reference's Beam<T,F> (Beam.h:69-114) and only observes top() and
// prune(); every line of search logic executed is the reference's.
#include <cstring>
#include <string>
#include <vector>

#include "BeamSearch.h"
#include "Forward.h"

namespace {

struct SpyLog {
  double top_score = 0;
  int top_depth = 0;
  long n_prune = 0;
  std::vector<double> step_top;   // score of elements[0] after each prune
  std::vector<int> step_depth;    // depth of elements[0] after each prune
  bool trace = false;
};
SpyLog g_spy;

enum ScoreKind { LAST = 0, MAX = 1, MAXSYM = 2 };

template <class T, class F, int K>
class SpyBeam : public Beam<T, F> {
 public:
  SpyBeam(int w) : Beam<T, F>(w) {}
  static double score(T n) {
    if (K == LAST) return n->last_probability();
    if (K == MAX) return n->max_probability();
    return n->max_probability_sym();
  }
  void prune() {
    Beam<T, F>::prune();
    g_spy.n_prune++;
    if (g_spy.trace && !this->elements.empty()) {
      g_spy.step_top.push_back(score(this->elements[0]));
      g_spy.step_depth.push_back(this->elements[0]->depth);
    }
  }
  T top() {
    T n = Beam<T, F>::top();
    g_spy.top_score = score(n);
    g_spy.top_depth = n->depth;
    return n;
  }
};

// 2D node types only define max_probability*(); 1D only last_probability().
// Guard the unused branches at compile time.
template <class T, class F>
class SpyBeam1D : public Beam<T, F> {
 public:
  SpyBeam1D(int w) : Beam<T, F>(w) {}
  void prune() {
    Beam<T, F>::prune();
    g_spy.n_prune++;
    if (g_spy.trace && !this->elements.empty()) {
      g_spy.step_top.push_back(this->elements[0]->last_probability());
      g_spy.step_depth.push_back(this->elements[0]->depth);
    }
  }
  T top() {
    T n = Beam<T, F>::top();
    g_spy.top_score = n->last_probability();
    g_spy.top_depth = n->depth;
    return n;
  }
};

std::vector<double*> row_ptrs(const double* y, int T, int S) {
  std::vector<double*> p(T > 0 ? T : 1);
  for (int t = 0; t < T; ++t) p[t] = const_cast<double*>(y) + (size_t)t * S;
  return p;
}

int put(const std::string& s, char* out, int cap) {
  // the reference label starts with alphabet[root->last], i.e. the string
  // terminator read past the alphabet (decoding_cpp.pyx:134 strips '\0')
  size_t b = 0;
  while (b < s.size() && s[b] == '\0') ++b;
  int n = (int)(s.size() - b);
  if (n + 1 > cap) return -n - 1;
  memcpy(out, s.data() + b, n);
  out[n] = 0;
  return n;
}

}  // namespace

extern "C" {

void ref_set_trace(int on) {
  g_spy.trace = on != 0;
  g_spy.step_top.clear();
  g_spy.step_depth.clear();
  g_spy.n_prune = 0;
}
long ref_trace_len() { return (long)g_spy.step_top.size(); }
void ref_trace_get(double* top, int* depth) {
  for (size_t i = 0; i < g_spy.step_top.size(); ++i) {
    top[i] = g_spy.step_top[i];
    depth[i] = g_spy.step_depth[i];
  }
}

// stock entry point, exactly what decoding_cpp.pyx:100 calls
int ref_beam_search_stock(const double* y, int T, int S, int W, const char* model, char* out, int cap) {
  auto p = row_ptrs(y, T, S);
  std::string alphabet("ACGT", S - 1 > 4 ? 4 : S - 1);
  if (std::string(model) == "ctc_flipflop") alphabet = std::string("ACGT", S / 2);
  std::string s = beam_search(p.data(), T, alphabet, W, std::string(model));
  return put(s, out, cap);
}

// stock entry points, exactly what decoding_cpp.pyx:130-133 calls
int ref_beam_search_2d_stock(const double* y1, const double* y2, int U, int V, int S, const int* env, int W,
                             const char* model, const char* method, char* out, int cap) {
  auto p1 = row_ptrs(y1, U, S);
  auto p2 = row_ptrs(y2, V, S);
  std::string alphabet("ACGT", S - 1 > 4 ? 4 : S - 1);
  if (std::string(model) == "ctc_flipflop") alphabet = std::string("ACGT", S / 2);
  std::string s;
  if (env) {
    std::vector<int*> e(U);
    for (int u = 0; u < U; ++u) e[u] = const_cast<int*>(env) + 2 * (size_t)u;
    s = beam_search(p1.data(), p2.data(), U, V, alphabet, e.data(), W, std::string(model), std::string(method));
  } else {
    s = beam_search(p1.data(), p2.data(), U, V, alphabet, W, std::string(model), std::string(method));
  }
  return put(s, out, cap);
}

// same template functions, spy beam -> also returns the ranking score of the returned node
int ref_beam_search(const double* y, int T, int S, int W, const char* model_, char* out, int cap, double* score) {
  auto p = row_ptrs(y, T, S);
  std::string model(model_), alphabet("ACGT", S - 1);
  std::string s;
  if (model == "ctc") {
    s = beam_search_<PoreOverPrefixTree, SpyBeam1D<PoreOverNode*, node_greater<PoreOverNode*>>>(p.data(), T, alphabet, W);
  } else if (model == "ctc_merge_repeats") {
    s = beam_search_<BonitoPrefixTree, SpyBeam1D<BonitoNode*, node_greater<BonitoNode*>>>(p.data(), T, alphabet, W);
  } else {
    return -1000000;
  }
  if (score) *score = g_spy.top_score;
  return put(s, out, cap);
}

int ref_beam_search_2d(const double* y1, const double* y2, int U, int V, int S, const int* env, int W,
                       const char* model_, const char* method_, char* out, int cap, double* score) {
  auto p1 = row_ptrs(y1, U, S);
  auto p2 = row_ptrs(y2, V, S);
  std::string model(model_), method(method_), alphabet("ACGT", S - 1);
  std::vector<int*> e;
  if (env) {
    e.resize(U);
    for (int u = 0; u < U; ++u) e[u] = const_cast<int*>(env) + 2 * (size_t)u;
  }
  std::string s;
  bool ctc = model == "ctc", bon = model == "ctc_merge_repeats";
  if (!ctc && !bon) return -1000000;
  if (method == "row") {
    typedef SpyBeam<PoreOverNode2D*, node_greater_max<PoreOverNode2D*>, MAX> BP;
    typedef SpyBeam<BonitoNode2D*, node_greater_max<BonitoNode2D*>, MAX> BB;
    if (env) {
      s = ctc ? beam_search_2d_by_row<PoreOverPrefixTree2D, BP>(p1.data(), p2.data(), e.data(), U, V, alphabet, W)
              : beam_search_2d_by_row<BonitoPrefixTree2D, BB>(p1.data(), p2.data(), e.data(), U, V, alphabet, W);
    } else {
      s = ctc ? beam_search_2d_by_row<PoreOverPrefixTree2D, BP>(p1.data(), p2.data(), U, V, alphabet, W)
              : beam_search_2d_by_row<BonitoPrefixTree2D, BB>(p1.data(), p2.data(), U, V, alphabet, W);
    }
  } else if (method == "row_col") {
    if (!env) return -1000001;
    typedef SpyBeam<PoreOverNode2D*, node_greater_max_sym<PoreOverNode2D*>, MAXSYM> BP;
    typedef SpyBeam<BonitoNode2D*, node_greater_max_sym<BonitoNode2D*>, MAXSYM> BB;
    s = ctc ? beam_search_2d_by_row_col<PoreOverPrefixTree2D, BP>(p1.data(), p2.data(), e.data(), U, V, alphabet, W)
            : beam_search_2d_by_row_col<BonitoPrefixTree2D, BB>(p1.data(), p2.data(), e.data(), U, V, alphabet, W);
  } else {
    return -1000002;
  }
  if (score) *score = g_spy.top_score;
  return put(s, out, cap);
}

// decoding_cpp.pyx:49-65 -> PrefixTree.h:751
double ref_forward(const double* y, int T, int S, const char* label, const char* model) {
  auto p = row_ptrs(y, T, S);
  std::string alphabet("ACGT", S - 1);
  return forward(p.data(), T, std::string(label), alphabet, std::string(model));
}

// decoding_cpp.pyx:69-84 -> Forward.h:14 (prints "Mapping label" on stdout, as the reference does)
int ref_viterbi_acceptor(const double* y, int T, int S, int band, const char* label, signed char* path_out) {
  auto p = row_ptrs(y, T, S);
  std::string alphabet("ACGT", S - 1);
  std::string path = viterbi_acceptor_poreover(p.data(), T, band, std::string(label), alphabet);
  for (int t = 0; t < T && t < (int)path.size(); ++t) path_out[t] = (signed char)(path[t] - '0');
  return 0;
}

}  // extern "C"
