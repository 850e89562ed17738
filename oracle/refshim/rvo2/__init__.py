"""Stand-in for the third-party ``rvo2`` module (sybrenstuvel/Python-RVO2), ORACLE ONLY.

Exposes the nine ``PyRVOSimulator`` methods the reference calls (orca.py:95-129,
crowd_sim.py:231-253) on top of the C restatement in ``oracle/crowdnav_oracle.c``.  Python floats are
rounded to float32 on entry (Cython ``float`` arguments) and returned as float32 values widened to
Python floats, exactly like the Cython binding.  Test infrastructure -- never imported by the product.
"""
import ctypes as C

import numpy as np

import oracle as _orc


class PyRVOSimulator(object):
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed,
                 velocity=(0, 0)):
        self._dt = np.float32(timeStep)
        self._defaults = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)
        self._pos, self._vel, self._pref = [], [], []
        self._radius, self._max_speed, self._nd, self._mn, self._th = [], [], [], [], []

    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None, timeHorizonObst=None,
                 radius=None, maxSpeed=None, velocity=None):
        d = self._defaults
        self._pos.append([np.float32(pos[0]), np.float32(pos[1])])
        v = d[6] if velocity is None else velocity
        self._vel.append([np.float32(v[0]), np.float32(v[1])])
        self._pref.append([np.float32(0), np.float32(0)])
        self._radius.append(np.float32(d[4] if radius is None else radius))
        self._max_speed.append(np.float32(d[5] if maxSpeed is None else maxSpeed))
        self._nd.append(np.float32(d[0] if neighborDist is None else neighborDist))
        self._mn.append(int(d[1] if maxNeighbors is None else maxNeighbors))
        self._th.append(np.float32(d[2] if timeHorizon is None else timeHorizon))
        return len(self._pos) - 1

    def getNumAgents(self):
        return len(self._pos)

    def setAgentPosition(self, i, pos):
        self._pos[i] = [np.float32(pos[0]), np.float32(pos[1])]

    def setAgentVelocity(self, i, vel):
        self._vel[i] = [np.float32(vel[0]), np.float32(vel[1])]

    def setAgentPrefVelocity(self, i, vel):
        self._pref[i] = [np.float32(vel[0]), np.float32(vel[1])]

    def getAgentVelocity(self, i):
        return float(self._vel[i][0]), float(self._vel[i][1])

    def getAgentPosition(self, i):
        return float(self._pos[i][0]), float(self._pos[i][1])

    def doStep(self):
        n = len(self._pos)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pos, vel, pref = f(self._pos), f(self._vel), f(self._pref)
        px, py, vx, vy = f(pos[:, 0]), f(pos[:, 1]), f(vel[:, 0]), f(vel[:, 1])
        prx, pry = f(pref[:, 0]), f(pref[:, 1])
        rad, ms, nd, th = f(self._radius), f(self._max_speed), f(self._nd), f(self._th)
        mn = np.ascontiguousarray(self._mn, dtype=np.int32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        _orc.lib().orc_rvo_do_step(n, fp(px), fp(py), fp(vx), fp(vy), fp(rad), fp(ms), fp(prx), fp(pry), fp(nd),
                                   mn.ctypes.data_as(C.POINTER(C.c_int)), fp(th), self._dt)
        self._pos = [[px[i], py[i]] for i in range(n)]
        self._vel = [[vx[i], vy[i]] for i in range(n)]
