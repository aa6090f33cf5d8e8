// fused 1-D by-tree kernels, float instantiations (see wx_tree1d.inl)
#define WX_TR_TYPE float
#include "wx_tree1d.inl"
