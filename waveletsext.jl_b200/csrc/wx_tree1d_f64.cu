// fused 1-D by-tree kernels, double instantiations (see wx_tree1d.inl)
#define WX_TR_TYPE double
#include "wx_tree1d.inl"
