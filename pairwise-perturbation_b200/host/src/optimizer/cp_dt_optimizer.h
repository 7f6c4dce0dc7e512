// cp_dt_optimizer.h -- standard dimension-tree ALS on a caterpillar tree over the local indices 0..N-2 of the
// rotated mode list; two first contractions per sweep, step() returns 0.5
// (reference: src/optimizer/cp_dt_optimizer.{h,cxx}).
#ifndef PPX_HOST_CP_DT_OPTIMIZER_H__
#define PPX_HOST_CP_DT_OPTIMIZER_H__

#include "cp_als_optimizer.h"

template <typename dtype>
class CPDTOptimizer : public CPOptimizer<dtype> {
public:
  CPDTOptimizer(int order, int r, World &dw) : CPOptimizer<dtype>(order, r, dw) {
    Construct_Dimension_Tree();
    indexes = vector<int>(order - 1, 0);
    for (size_t i = 0; i < indexes.size(); i++) indexes[i] = (int)i;
    indexes1 = indexes;
    indexes2 = indexes1;
    left_index = order - 1;
    left_index1 = left_index;
    left_index2 = (left_index + order - 1) % order;
    update_indexes(indexes2, left_index2);
    special_index = 0;
    first_subtree = true;
  }
  ~CPDTOptimizer() {}

  // cp_dt_optimizer.cxx:188-238
  double step() {
    if (first_subtree) {
      indexes = indexes1;
      left_index = left_index1;
    } else {
      indexes = indexes2;
      left_index = left_index2;
    }
    mttkrp_map.clear();
    mttkrp_map_init(left_index);
    for (int i = 0; i < (int)indexes.size(); i++) {
      if (first_subtree && i < special_index) continue;
      if (!first_subtree && i > special_index) break;
      update_leaf(i);
    }
    first_subtree = !first_subtree;
    return 0.5;
  }

  void update_left_index() { left_index = (left_index + this->order - 1) % this->order; }

  // modes left+1..N-1, 0..left-1  (cp_dt_optimizer.cxx:52-66)
  void update_indexes(vector<int> &idx, int left) {
    int j = 0;
    for (int i = left + 1; i < this->order; i++) idx[j++] = i;
    for (int i = 0; i < left; i++) idx[j++] = i;
  }

  void Construct_Dimension_Tree() {
    vector<int> top_node(this->order - 1);
    for (size_t i = 0; i < top_node.size(); i++) top_node[i] = (int)i;
    Construct_Subtree(top_node);
  }

  // cp_dt_optimizer.cxx:78-100: left child drops the last local index
  void Construct_Subtree(vector<int> top_node) {
    Right_Subtree(top_node);
    vector<int> child(top_node.begin(), top_node.end() - 1);
    link(child, top_node, top_node.back());
    if (child.size() > 1) Construct_Subtree(child);
  }

  // cp_dt_optimizer.cxx:102-124: right child drops the second-to-last local index
  void Right_Subtree(vector<int> top_node) {
    vector<int> child(top_node.begin(), top_node.end() - 1);
    child.back() = top_node.back();
    link(child, top_node, top_node[top_node.size() - 2]);
    if (child.size() > 1) Right_Subtree(child);
  }

  // root of the tree: V x W[left]  -- THE first contraction (cp_dt_optimizer.cxx:127-160)
  void mttkrp_map_init(int left) {
    World &dw = *this->world;
    const int order = this->order;
    string modes;
    for (int i = 0; i < order; i++) modes.push_back((char)('a' + i));
    Tensor<dtype> root = contract_mode(*this->V, modes, false, (char)('a' + left), this->W[left], dw);
    // `root` keeps its modes in increasing order (0..left-1, left+1..N-1, rank); the reference stores them rotated
    // (left+1..N-1, 0..left-1).  Remember which global mode each stored axis holds instead of permuting the data.
    vector<int> axes;
    for (int i = 0; i < order; i++)
      if (i != left) axes.push_back(i);
    string top;
    vector<int> ids(order - 1);
    for (int i = 0; i < order - 1; i++) ids[i] = i;
    vec2str(ids, top);
    mttkrp_map[top] = std::move(root);
    axes_map[top] = axes;
  }

  // cp_dt_optimizer.cxx:163-186
  void mttkrp_map_DT(string index) {
    World &dw = *this->world;
    const string par = parent[index];
    if (mttkrp_map.find(par) == mttkrp_map.end()) mttkrp_map_DT(par);
    const int local = contract_index[index][0] - 'a';
    const int gmode = indexes[local];
    const vector<int> &pax = axes_map[par];
    string modes;
    for (int m : pax) modes.push_back((char)('a' + m));
    mttkrp_map[index] = contract_mode(mttkrp_map[par], modes, true, (char)('a' + gmode), this->W[gmode], dw);
    vector<int> ax;
    for (int m : pax)
      if (m != gmode) ax.push_back(m);
    axes_map[index] = ax;
  }

  map<string, Tensor<dtype>> mttkrp_map;
  map<string, string> parent;
  map<string, string> contract_index;
  bool first_subtree;
  vector<int> indexes, indexes1, indexes2;
  int left_index, left_index1, left_index2;
  int special_index;

protected:
  map<string, vector<int>> axes_map;  // global modes held by each cached tensor, in storage order
  void link(const vector<int> &child, const vector<int> &top, int mat) {
    string c, t, m;
    vec2str(child, c);
    vec2str(top, t);
    vec2str(vector<int>{mat}, m);
    parent[c] = t;
    contract_index[c] = m;
  }
  // MTTKRP of local index i from the tree (a copy: the solve kernels may overwrite it)
  Matrix<dtype> leaf(int i) {
    string key;
    vec2str(vector<int>{i}, key);
    if (mttkrp_map.find(key) == mttkrp_map.end()) mttkrp_map_DT(key);
    return Matrix<dtype>(mttkrp_map[key]);
  }
  void update_leaf(int i) {
    Matrix<dtype> M = leaf(i);
    this->solve_mode(indexes[i], M);
  }
};

#endif
