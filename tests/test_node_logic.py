"""Host logic of the path's caller: the keyframe state machine of rmd::DepthmapNode
(src/depthmap_node.cpp:88-183) against a scripted stand-in for rmd::Depthmap (CPU), and several live
keyframes fed through rmd_seeds_update_many against the same keyframes updated one by one (GPU, bit-exact)."""
import numpy as np
import pytest

from rpg_open_remode_b200.api import SE3
from rpg_open_remode_b200 import node


class ScriptedDepthmap:
    """Records the calls the node makes; convergence and distance follow a script per update."""

    def __init__(self, script, accept_reference=True):
        self.script, self.accept, self.calls, self.i = list(script), accept_reference, [], -1

    def setReferenceImage(self, img, T, dmin, dmax):
        self.calls.append(("setReferenceImage", T.data.copy(), dmin, dmax))
        return self.accept

    def update(self, img, T):
        self.i += 1
        self.calls.append(("update", T.data.copy()))

    def getConvergedPercentage(self):
        return self.script[self.i][0]

    def getDistFromRef(self):
        return self.script[self.i][1]

    def downloadDenoisedDepthmap(self, lam, iters):
        self.calls.append(("denoise", lam, iters))

    def downloadConvergenceMap(self):
        self.calls.append(("downloadConvergenceMap",))


def _names(calls):
    return [c[0] for c in calls]


def test_state_machine_rekeyframes_on_convergence_or_distance():
    published = []
    dm = ScriptedDepthmap([(2.0, 0.1), (11.0, 0.1),        # 11 % > 10 %: keyframe finished
                           (0.0, 0.2), (3.0, 0.6),         # 0.6 m > 0.5 m: keyframe finished
                           (10.0, 0.5)])                   # exactly at both thresholds: NOT finished (strict >)
    n = node.DepthmapNode(dm, publish_conv_every_n=100, publisher=lambda what, d: published.append(what))
    T = SE3(0.9238795, 0.0, 0.3826834, 0.0, 1.0, 2.0, 3.0)     # T_world_curr as the message carries it
    img = np.zeros((4, 4), np.uint8)
    assert n.state_ == node.TAKE_REFERENCE_FRAME
    for _ in range(8):
        n.denseInputCallback(img, T, 0.5, 2.5)
    assert _names(dm.calls) == ["setReferenceImage", "update", "update", "denoise", "downloadConvergenceMap",
                                "setReferenceImage", "update", "update", "denoise", "downloadConvergenceMap",
                                "setReferenceImage", "update"]
    assert published == ["depthmap_and_pointcloud"] * 2 and n.state_ == node.UPDATE
    # the Depthmap receives T_curr_world = T_world_curr.inv() (src/depthmap_node.cpp:128,142); denoise(0.5, 200) (:167)
    assert np.array_equal(dm.calls[0][1], T.inv().data) and dm.calls[0][2:] == (0.5, 2.5)
    assert dm.calls[3] == ("denoise", 0.5, 200)


def test_rejected_reference_keeps_waiting_and_convergence_map_cadence():
    published = []
    dm = ScriptedDepthmap([(0.0, 0.0)] * 20, accept_reference=False)
    n = node.DepthmapNode(dm, publish_conv_every_n=3, publisher=lambda what, d: published.append(what))
    img, T = np.zeros((4, 4), np.uint8), SE3()
    for _ in range(9):
        n.denseInputCallback(img, T, 1.0, 2.0)
    assert _names(dm.calls).count("setReferenceImage") == 9 and "update" not in _names(dm.calls)
    # publish_conv_every_n_ < num_msgs_ (:157): every 4th message with n = 3, counter reset after publishing
    assert published == ["convergence", "convergence"] and n.num_msgs_ == 1
    with pytest.raises(RuntimeError):
        node.DepthmapNode(None).denseInputCallback(img, T, 1.0, 2.0)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "uint8"])
def test_live_keyframes_updated_together_equal_one_by_one(qvga_sequence, dtype):
    import rpg_open_remode_b200 as rmd
    seq = qvga_sequence
    cam = rmd.PinholeCamera(*seq.camera)
    pick = (lambda f: f.image_u8) if dtype == "uint8" else (lambda f: f.image)
    starts = [0, 4, 9]                      # three keyframes taken at different frames of one stream
    ks = node.KeyframeSet(seq.width, seq.height, cam, n=len(starts))
    alone = [rmd.SeedMatrix(seq.width, seq.height, cam) for _ in starts]
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    for k in range(0, 30):
        f = seq.frame(k, want_depth=False)
        assert ks.update(pick(f), f.T_cam_world) == sum(k > s for s in starts)
        for i, s in enumerate(starts):
            if k > s:
                alone[i].update(pick(f), f.T_cam_world)
        if k in starts:
            i = starts.index(k)
            ks.setReferenceImage(i, pick(f), f.T_cam_world, dmin, dmax)
            alone[i].setReferenceImage(pick(f), f.T_cam_world, dmin, dmax)
    for i in range(len(starts)):
        for get in ("downloadConvergence", "downloadDepthmap", "downloadSigmaSq", "downloadA", "downloadB"):
            assert np.array_equal(getattr(ks.seeds[i], get)(), getattr(alone[i], get)()), (i, get)
        assert ks.seeds[i].getConvergedCount() == alone[i].getConvergedCount()
    assert ks.convergedPercentages().shape == (3,)
    ks.retire(1)
    f = seq.frame(30, want_depth=False)
    assert ks.update(pick(f), f.T_cam_world) == 2
