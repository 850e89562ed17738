"""Host-side scene generation with numpy's legacy MT19937 stream, exactly as CrowdSim.reset seeds and draws
it (crowd_sim/envs/crowd_sim.py:165-217,261-323).  Used for parity runs and for the single-env façade;
throughput runs generate scenes on the GPU (cn_env_reset, Philox).
"""
import numpy as np

# crowd_sim.py:68,282-283: counter_offset = {'train': val+test capacity, 'val': 0, 'test': val capacity}
COUNTER_OFFSET = {"train": 2000, "val": 0, "test": 1000}


def _norm(a, b):
    return float(np.linalg.norm((a, b)))


def _init_v(px, py, gx, gy, v_pref):
    """model_crowd_sim.py:186-192: goal direction scaled so that its larger component is v_pref."""
    vx, vy = gx - px, gy - py
    vmax = max(abs(vx), abs(vy))
    return v_pref * vx / vmax, v_pref * vy / vmax


# crowd_sim.py:113-115: number of humans of a 'mixed' scene, by cumulative probability over the sorted keys
MIXED_STATIC_NUM = {0: 0.05, 1: 0.2, 2: 0.2, 3: 0.3, 4: 0.1, 5: 0.15}
MIXED_DYNAMIC_NUM = {1: 0.3, 2: 0.3, 3: 0.2, 4: 0.1, 5: 0.1}


def generate_scene(phase, case, human_num=5, rule="circle_crossing", circle_radius=4.0, square_width=10.0,
                   human_radius=0.3, human_v_pref=1.0, discomfort_dist=0.2, robot_radius=0.3, robot_v_pref=1.0,
                   randomize_attributes=False, rs=None, init_velocity=False):
    """Agents (H+1, 8) float64 [px py vx vy gx gy radius v_pref]; agent 0 = robot (crowd_sim.py:284).
    randomize_attributes ([env] randomize_attributes): each human first draws v_pref ~ U(0.5, 1.5) and
    radius ~ U(0.3, 0.5) from the same stream (crowd_sim.py:167-168,190-191; agent.py:39-45).
    rule = 'mixed' (crowd_sim.py:111-161): the scene draws its own number of humans (the human_num argument is ignored, as
    in the reference) -- 20 % static scenes of 0-5 standing humans (goal = position; 0 humans = one dummy human parked at
    (0, -10)), else 1-5 moving humans, the first two circle-crossing, the rest square-crossing."""
    if rs is None:
        rs = np.random.RandomState(COUNTER_OFFSET[phase] + case)
    # rs given (ModelCrowdSim.reset never seeds, model_crowd_sim.py:293: pass the np.random module to draw from the global
    # stream like the reference); init_velocity: humans start moving towards their goal (gen_init_v, model_crowd_sim.py:186-192)
    robot = [0, -circle_radius, 0, 0, 0, circle_radius, robot_radius, robot_v_pref]
    base_radius, base_v_pref = human_radius, human_v_pref
    rules = None
    if rule == "mixed":
        static = bool(rs.random_sample() < 0.2)
        prob = rs.random_sample()
        for key, value in sorted((MIXED_STATIC_NUM if static else MIXED_DYNAMIC_NUM).items()):
            if prob - value <= 0:
                human_num = key
                break
            prob -= value
        if static:
            rows = [robot]
            if human_num == 0:
                rows.append([0, -10, 0, 0, 0, -10, base_radius, base_v_pref])
            for _ in range(human_num):
                sign = -1 if rs.random_sample() > 0.5 else 1
                while True:
                    px = rs.random_sample() * 4 * 0.5 * sign                  # width = 4
                    py = (rs.random_sample() - 0.5) * 8                       # height = 8
                    if not any(_norm(px - a[0], py - a[1]) < base_radius + a[6] + discomfort_dist for a in rows):
                        break
                rows.append([px, py, 0, 0, px, py, base_radius, base_v_pref])
            return np.array(rows, dtype=np.float64)
        rules = ["circle_crossing" if i < 2 else "square_crossing" for i in range(human_num)]
    agents = np.zeros((human_num + 1, 8))
    agents[0] = robot
    for i in range(1, human_num + 1):
        prev = agents[:i]
        human_radius, human_v_pref = base_radius, base_v_pref
        if rules is not None:
            rule = rules[i - 1]
        if randomize_attributes:
            human_v_pref = rs.uniform(0.5, 1.5)
            human_radius = rs.uniform(0.3, 0.5)
        if rule == "circle_crossing":
            while True:
                angle = rs.random_sample() * np.pi * 2
                px_noise = (rs.random_sample() - 0.5) * human_v_pref
                py_noise = (rs.random_sample() - 0.5) * human_v_pref
                px = circle_radius * np.cos(angle) + px_noise
                py = circle_radius * np.sin(angle) + py_noise
                collide = False
                for a in prev:
                    min_dist = human_radius + a[6] + discomfort_dist
                    if _norm(px - a[0], py - a[1]) < min_dist or _norm(px - a[4], py - a[5]) < min_dist:
                        collide = True
                        break
                if not collide:
                    break
            agents[i] = [px, py, 0, 0, -px, -py, human_radius, human_v_pref]
            if init_velocity:
                agents[i, 2:4] = _init_v(px, py, -px, -py, human_v_pref)
        elif rule == "square_crossing":
            sign = -1 if rs.random_sample() > 0.5 else 1
            while True:
                px = rs.random_sample() * square_width * 0.5 * sign
                py = (rs.random_sample() - 0.5) * square_width
                if not any(_norm(px - a[0], py - a[1]) < human_radius + a[6] + discomfort_dist for a in prev):
                    break
            while True:
                gx = rs.random_sample() * square_width * 0.5 * -sign
                gy = (rs.random_sample() - 0.5) * square_width
                if not any(_norm(gx - a[4], gy - a[5]) < human_radius + a[6] + discomfort_dist for a in prev):
                    break
            agents[i] = [px, py, 0, 0, gx, gy, human_radius, human_v_pref]
            if init_velocity:                  # model_crowd_sim.py:222: towards (-px, -py), not towards the square goal
                agents[i, 2:4] = _init_v(px, py, -px, -py, human_v_pref)
        else:
            raise ValueError("Rule doesn't exist")
    return agents


def generate_batch(phase, cases, **kw):
    """Scenes of many cases at once.  Plain circle / square crossing scenes seeded per case go through the native generator
    of the C ABI (cn_scenes_generate: the same MT19937 stream and arithmetic, bit-identical, ~1000x faster than the Python
    loop); 'mixed' scenes, the global-stream / initial-velocity variants of ModelCrowdSim stay in Python."""
    cases = [int(c) for c in cases]
    rule = kw.get("rule", "circle_crossing")
    if rule in ("circle_crossing", "square_crossing") and kw.get("rs") is None and not kw.get("init_velocity") and cases:
        import ctypes as C
        from . import _capi
        lib = _capi.load()
        H = int(kw.get("human_num", 5))
        seeds = np.array([COUNTER_OFFSET[phase] + c for c in cases], dtype=np.int64)
        out = np.empty((len(cases), H + 1, 8), np.float64)
        _capi.check(lib.cn_scenes_generate(len(cases), seeds.ctypes.data_as(C.c_void_p), H,
                                           _capi.CIRCLE_CROSSING if rule == "circle_crossing" else _capi.SQUARE_CROSSING,
                                           float(kw.get("circle_radius", 4.0)), float(kw.get("square_width", 10.0)),
                                           float(kw.get("human_radius", 0.3)), float(kw.get("human_v_pref", 1.0)),
                                           float(kw.get("discomfort_dist", 0.2)), float(kw.get("robot_radius", 0.3)),
                                           float(kw.get("robot_v_pref", 1.0)), int(bool(kw.get("randomize_attributes", False))),
                                           out.ctypes.data_as(C.c_void_p)))
        return out
    return np.stack([generate_scene(phase, c, **kw) for c in cases])
