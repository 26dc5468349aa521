"""The caller of the hot path: rmd::DepthmapNode's keyframe state machine
without ROS, and a set of several live keyframes fed by one frame stream.

  DepthmapNode   src/depthmap_node.cpp:36-183 + include/rmd/depthmap_node.h:30-62: TAKE_REFERENCE_FRAME /
                 UPDATE, re-keyframing when the converged percentage or the distance from the reference
                 passes its threshold, then denoise(0.5, 200) + convergence download + publish.
  KeyframeSet    what the reference cannot do (one rmd::Depthmap, so the map of a keyframe stops growing
                 the moment the next one starts): up to `n` keyframes stay live and every frame updates all of
                 them through rmd_seeds_update_many -- one upload, one fused kernel per keyframe, overlapping
                 on the GPU.

Host logic only: everything numeric happens behind rpg_open_remode_b200.api (the C-ABI).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np

from .api import SE3, Depthmap, SeedMatrix

UPDATE, TAKE_REFERENCE_FRAME = 0, 1   # rmd::ProcessingStates::State, include/rmd/depthmap_node.h:32-36


class DepthmapNode:
    """rmd::DepthmapNode::denseInputCallback with the ROS plumbing removed.  `publisher` receives
    ("depthmap_and_pointcloud", depthmap) after a keyframe is finished (denoiseAndPublishResults,
    src/depthmap_node.cpp:165-173) and ("convergence", depthmap) every publish_conv_every_n messages
    (:157-161, :175-183)."""

    def __init__(self, depthmap: Depthmap, ref_compl_perc: float = 10.0, max_dist_from_ref: float = 0.5,
                 publish_conv_every_n: int = 10, publisher: Optional[Callable] = None):
        self.depthmap_ = depthmap
        self.state_ = TAKE_REFERENCE_FRAME                      # src/depthmap_node.cpp:35
        self.ref_compl_perc_ = float(ref_compl_perc)            # :81, default 10.0
        self.max_dist_from_ref_ = float(max_dist_from_ref)      # :82, default 0.5
        self.publish_conv_every_n_ = int(publish_conv_every_n)  # :83, default 10
        self.num_msgs_ = 0
        self.publisher_ = publisher

    def denseInputCallback(self, img_8uc1, T_world_curr: SE3, min_depth: float, max_depth: float) -> None:
        self.num_msgs_ += 1                                                       # :92
        if self.depthmap_ is None:
            raise RuntimeError("depthmap not initialized")                        # :93-97
        T_curr_world = T_world_curr.inv()                                         # :128, :142
        if self.state_ == TAKE_REFERENCE_FRAME:
            if self.depthmap_.setReferenceImage(img_8uc1, T_curr_world, min_depth, max_depth):
                self.state_ = UPDATE                                              # :126-135
        elif self.state_ == UPDATE:
            self.depthmap_.update(img_8uc1, T_curr_world)                         # :142
            perc_conv = self.depthmap_.getConvergedPercentage()                   # :143
            dist_from_ref = self.depthmap_.getDistFromRef()                       # :144
            if perc_conv > self.ref_compl_perc_ or dist_from_ref > self.max_dist_from_ref_:   # :146
                self.state_ = TAKE_REFERENCE_FRAME
                self.denoiseAndPublishResults()
        if self.publish_conv_every_n_ < self.num_msgs_:                           # :157
            self.publishConvergenceMap()
            self.num_msgs_ = 0

    def denoiseAndPublishResults(self) -> None:
        self.depthmap_.downloadDenoisedDepthmap(0.5, 200)                         # :167
        self.depthmap_.downloadConvergenceMap()                                   # :168
        if self.publisher_:
            self.publisher_("depthmap_and_pointcloud", self.depthmap_)

    def publishConvergenceMap(self) -> None:
        self.depthmap_.downloadConvergenceMap()                                   # :177
        if self.publisher_:
            self.publisher_("convergence", self.depthmap_)


class KeyframeSet:
    """Up to `n` live keyframes of one camera; update() feeds a frame to all of them at once."""

    def __init__(self, width: int, height: int, camera, n: int, patch_side: int = 5, device: int = -1):
        self.seeds: List[SeedMatrix] = [SeedMatrix(width, height, camera, patch_side, device) for _ in range(n)]
        self.live: List[bool] = [False] * n

    def setReferenceImage(self, slot: int, img, T_curr_world, min_depth: float, max_depth: float) -> None:
        self.seeds[slot].setReferenceImage(img, T_curr_world, min_depth, max_depth)
        self.live[slot] = True

    def retire(self, slot: int) -> None:
        self.live[slot] = False

    def update(self, img, T_curr_world) -> int:
        """Returns the number of keyframes the frame went to."""
        live = [s for s, on in zip(self.seeds, self.live) if on]
        if live:
            SeedMatrix.updateMany(live, img, T_curr_world)
        return len(live)

    def convergedPercentages(self) -> np.ndarray:
        return np.array([100.0 * s.getConvergedCount() / (s.width_ * s.height_) if on else np.nan
                         for s, on in zip(self.seeds, self.live)], np.float32)
