"""Host-side mirror of the belief filter of the reference's DummySimulator
(/root/reference/dummy_simulator/src/dummy_simulator.cpp:671-773) over the C
ABI.  The reference keeps ONE belief and updates it on the CPU once per
control message; here any number of independent beliefs (Monte-Carlo runs of
a planner, many simulated robots) are filtered in one call on the GPU, each
with its own action / measurement, with the reference's bits."""
import ctypes

import numpy as np

from . import _lib


class DummySimulator:
    def __init__(self, grid_map):
        grid_map = np.ascontiguousarray(grid_map, dtype=np.uint8)
        self.map_height, self.map_width = grid_map.shape
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.pp2d_sim_create(self.map_height, self.map_width,
                                             grid_map.ctypes.data, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pp2d_sim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _beliefs(self, beliefs):
        b = np.array(beliefs, dtype=np.float32, order="C", copy=True)
        return b.reshape(-1, self.map_height * self.map_width)

    def updateBelief(self, beliefs, action=None, measurement=None):
        """updateBelief(u) when `action` is given (one value or one per belief),
        updateBelief(meas) when `measurement` is given ([4] or [n][4]); both:
        controlCallback's order, action first.  Returns the new beliefs."""
        b = self._beliefs(beliefs)
        n = b.shape[0]
        a = m = None
        if action is not None:
            a = np.ascontiguousarray(np.broadcast_to(action, (n,)), dtype=np.uint8)
        if measurement is not None:
            m = np.ascontiguousarray(np.broadcast_to(measurement, (n, 4)), dtype=np.uint8)
        if a is not None and m is not None:
            rc = self._lib.pp2d_sim_step(self._h, b.ctypes.data, n, a.ctypes.data, m.ctypes.data)
        elif a is not None:
            rc = self._lib.pp2d_sim_update_belief_action(self._h, b.ctypes.data, n, a.ctypes.data)
        elif m is not None:
            rc = self._lib.pp2d_sim_update_belief_measurement(self._h, b.ctypes.data, n,
                                                              m.ctypes.data)
        else:
            raise ValueError("give an action, a measurement, or both")
        _lib.check(rc)
        return b
