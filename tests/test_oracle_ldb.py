"""CPU checks of the oracle's restatement of the LDB object (LDB.jl:186-245, ldb/ldb_measures.jl:427-479) on hand-computed cases.  No GPU."""
import numpy as np


def test_fisher_power_hand_computed(O):
    # two classes, two coefficients: class means (2, 12) / (2, 2), variances (2, 8) / (0, 8), p = (1/2, 1/2)
    c = np.array([[1., 2.], [3., 2.], [10., 0.], [14., 4.]]); y = ["a", "a", "b", "b"]
    power, order = O.discriminant_power_fisher(c, y)
    E = np.array([[2., 12.], [2., 2.]]); V = np.array([[2., 8.], [0., 8.]]); p = np.array([.5, .5])
    Ea = E.mean(1, keepdims=True)
    want = (((E - Ea * E) ** 2) @ p) / (V @ p)                       # the reference's formula, (Eα .* Eαᵢ) included
    assert np.allclose(power, want, rtol=1e-15) and list(order) == [0, 1]
    assert power[0] == 532.8 and power[1] == 1.0


def test_topk_costs_and_basis_power(O):
    DM = np.array([[5., 1., 4., 2., 3., 9., 0., 7.],                # level 0: one node of 8
                   [8., 1., 1., 1., 2., 2., 2., 2.],                # level 1: two nodes of 4
                   [3., 3., 0., 0., 9., 1., 1., 1.]])               # level 2: four nodes of 2
    c_all = O.ldb_costs_topk(DM, 8)
    assert np.array_equal(c_all, O.ldb_costs(DM))
    assert list(c_all) == [31., 11., 8., 6., 0., 10., 2.]
    c2 = O.ldb_costs_topk(DM, 2)                                      # two largest per node (nodes of 2 keep their sum)
    assert list(c2) == [16., 9., 4., 6., 0., 10., 2.]
    tree = O.tree_select(c2.copy(), 8, None, "max")                   # 16 vs children 9 + max(4, 10 + 2): 9 + 12 = 21 > 16
    assert list(tree[:3]) == [True, False, True]
    power, order = O.discriminant_power_basis(DM, tree)
    assert list(power) == [8., 1., 1., 1., 9., 1., 1., 1.]            # level-1 left node, level-2 right pair
    assert list(order[:2]) == [4, 0] and list(order[2:]) == [1, 2, 3, 5, 6, 7]      # stable among ties (sortperm rev=true)
