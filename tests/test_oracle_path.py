"""CPU checks of the path oracle (oracle/path.cpp): hand-derived known answers for the intent-mode rules."""
import struct

import numpy as np

import oracle


def _grid(W, H, h=None):
    n = W * H
    height = np.zeros(n, np.float32) if h is None else np.asarray(h, np.float32).reshape(n)
    pos = np.zeros((n, 3), np.float32)
    pos[:, 0] = np.tile(np.arange(W), H)
    pos[:, 1] = height
    pos[:, 2] = np.repeat(np.arange(H), W)
    conn = np.ones((n, 8), np.float32)
    return height, pos, conn


def test_flat_grid_costs_are_hop_counts_along_the_flat_index():
    """every edge weighs 1 (connections == 1, flat height): cost = graph distance in Scene::neighbors' graph, where
    px-1 / px+1 also join the end of a row to the start of the next (scene.rs:136-137)"""
    W, H = 8, 6
    height, pos, conn = _grid(W, H)
    balls = np.array([[3, 2], [3, 2], [3, 2]], np.int32)
    r = oracle.path_modify(height, pos, balls, conn, W, H)
    t = 3 + 2 * W
    assert r["cost"][t] == 0 and r["pred"][t] == -2
    assert r["cost"][t + 1] == 1 and r["cost"][t - W] == 1 and r["cost"][t + W + 1] == 2
    # wrap: node 7 (end of row 0) and node 8 (start of row 1) are neighbours
    assert abs(r["cost"][7] - r["cost"][8]) <= 1
    # the start node W*H - H/2 walks to the target in cost[start] unit steps
    start = W * H - H // 2
    assert r["n"] == int(r["cost"][start])
    assert np.all(r["directions"][:, 0] == 1.0) and r["directions"][0, 1] == 0.0


def test_height_steps_and_absent_edges():
    W, H = 6, 4
    h = np.zeros((H, W), np.float32)
    h[:, 3] = 5.0                       # a wall in column 3: stepping onto it costs 1 + 5, and so does stepping off
    height, pos, conn = _grid(W, H, h)
    balls = np.array([[0, 0]] * 3, np.int32)
    r = oracle.path_modify(height, pos, balls, conn, W, H)
    assert r["cost"][2] == 2.0 and r["cost"][3] == 2.0 + 1.0 + 5.0
    # the far side of the wall is reached around it: px+1 of a row's last pixel is the next row's first (scene.rs:137)
    assert r["cost"][6] == 1.0 and r["cost"][5] == 2.0 and r["cost"][4] == 3.0
    conn[:] = -1.0                      # the shaders' "no neighbour" marker removes every edge
    r = oracle.path_modify(height, pos, balls, conn, W, H)
    assert r["directions"] is None and (r["cost"][1:] > 3e38).all()


def test_serialize_is_big_endian():
    s = oracle.path_serialize(0x0102030405060708, np.array([[1.0, -2.0]], np.float32))
    assert s == struct.pack(">Qff", 0x0102030405060708, 1.0, -2.0)      # path.rs:17-21
    assert oracle.lib().tod_oracle_path_literal_panics() == 1
