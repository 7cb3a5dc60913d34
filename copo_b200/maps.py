"""Lane-geometry tables for the five MARL maps, packed as one u32 "map blob".

The reference gets its maps from MetaDrive 0.2.5 (`metadrive.envs.marl_envs.MultiAgent*Env`,
imported at copo_code/copo/torch_copo/train_copo.py:1-2); MetaDrive is not vendored, so the
geometry below is this repo's own specification (DESIGN.md "Simulator spec").  A map is
*input data*: the builder runs on the host in float64, the tables are frozen to float32 and the
same blob is consumed by the CUDA step kernel (staged into shared memory with one bulk copy)
and by the oracle (tests hand it the identical arrays).

Blob layout (u32 words, all offsets in words):
  header[16] : magic, n_seg, n_route, n_spawn, off_seg, off_route, off_spawn, off_ray,
               total_words, route_stride, n_ray, base_obs_dim, n_side, reserved...
  seg[n_seg][12]   f32: x0 y0 h0 len kappa wl wr cos(h0) sin(h0) cx cy r
  route[n_route][8]   : i32 nseg, i32 seg[6], f32 total_len
  spawn[n_spawn][12]  : f32 x y h cos sin s0, i32 seg0, i32 n_routes, i32 route[4]
  ray[n_ray][2]    f32: cos, sin of the laser directions in the ego frame
"""
import math

import numpy as np

MAGIC = 0xB200C0F0
HEADER_WORDS = 16
SEG_WORDS = 12
ROUTE_WORDS = 8
ROUTE_MAX_SEGS = 6
SPAWN_WORDS = 12
SPAWN_MAX_ROUTES = 4
NUM_LASERS = 72
LANE_WIDTH = 3.5


class _Builder:
    """Chains straight / arc centre-lines; de-duplicates shared segments."""

    def __init__(self):
        self.segs = []
        self.seg_index = {}
        self.routes = []
        self.spawns = []

    # -- segments ---------------------------------------------------------
    def _add_seg(self, x0, y0, h0, length, kappa, wl, wr):
        key = tuple(round(v, 6) for v in (x0, y0, h0, length, kappa, wl, wr))
        if key in self.seg_index:
            return self.seg_index[key]
        c0, s0 = math.cos(h0), math.sin(h0)
        if kappa != 0.0:
            r = 1.0 / abs(kappa)
            sg = 1.0 if kappa > 0 else -1.0
            cx = x0 - s0 * r * sg
            cy = y0 + c0 * r * sg
        else:
            r, cx, cy = 0.0, 0.0, 0.0
        self.segs.append([x0, y0, h0, length, kappa, wl, wr, c0, s0, cx, cy, r])
        self.seg_index[key] = len(self.segs) - 1
        return len(self.segs) - 1

    @staticmethod
    def _end_pose(x0, y0, h0, length, kappa):
        if kappa == 0.0:
            return x0 + length * math.cos(h0), y0 + length * math.sin(h0), h0
        h1 = h0 + kappa * length
        x1 = x0 + (math.sin(h1) - math.sin(h0)) / kappa
        y1 = y0 - (math.cos(h1) - math.cos(h0)) / kappa
        return x1, y1, h1

    def chain(self, pose, pieces):
        """pieces: list of ("s", len, wl, wr) | ("a", radius, signed_angle, wl, wr). Returns seg ids."""
        x, y, h = pose
        ids = []
        for p in pieces:
            if p[0] == "s":
                _, length, wl, wr = p
                kappa = 0.0
            else:
                _, radius, ang, wl, wr = p
                kappa = (1.0 if ang > 0 else -1.0) / radius
                length = abs(ang) * radius
            ids.append(self._add_seg(x, y, _wrap(h), length, kappa, wl, wr))
            x, y, h = self._end_pose(x, y, h, length, kappa)
        return ids

    def add_route(self, seg_ids):
        assert 1 <= len(seg_ids) <= ROUTE_MAX_SEGS
        total = sum(self.segs[i][3] for i in seg_ids)
        self.routes.append((list(seg_ids), total))
        return len(self.routes) - 1

    def add_spawn(self, seg0, s0, route_ids):
        x0, y0, h0 = self.segs[seg0][0:3]
        assert self.segs[seg0][4] == 0.0, "spawn places sit on straight lanes"
        x = x0 + s0 * math.cos(h0)
        y = y0 + s0 * math.sin(h0)
        assert 1 <= len(route_ids) <= SPAWN_MAX_ROUTES
        self.spawns.append((x, y, h0, s0, seg0, list(route_ids)))

    # -- packing ----------------------------------------------------------
    def pack(self, name, base_obs_dim=91, n_side=0):
        n_seg, n_route, n_spawn = len(self.segs), len(self.routes), len(self.spawns)
        off_seg = HEADER_WORDS
        off_route = off_seg + n_seg * SEG_WORDS
        off_spawn = off_route + n_route * ROUTE_WORDS
        off_ray = off_spawn + n_spawn * SPAWN_WORDS
        total = off_ray + NUM_LASERS * 2
        total = (total + 3) // 4 * 4  # 16-byte multiple for the bulk copy
        blob = np.zeros(total, dtype=np.uint32)
        f = blob.view(np.float32)
        i = blob.view(np.int32)
        blob[0:13] = [MAGIC, n_seg, n_route, n_spawn, off_seg, off_route, off_spawn, off_ray, total,
                      ROUTE_WORDS, NUM_LASERS, base_obs_dim, n_side]
        seg = np.asarray(self.segs, dtype=np.float64).astype(np.float32)
        f[off_seg:off_route] = seg.reshape(-1)
        for r, (ids, tot) in enumerate(self.routes):
            o = off_route + r * ROUTE_WORDS
            i[o] = len(ids)
            for k, sid in enumerate(ids):
                i[o + 1 + k] = sid
            f[o + 7] = np.float32(tot)
        for p, (x, y, h, s0, seg0, rids) in enumerate(self.spawns):
            o = off_spawn + p * SPAWN_WORDS
            f[o:o + 6] = np.asarray([x, y, h, math.cos(h), math.sin(h), s0], dtype=np.float64).astype(np.float32)
            i[o + 6] = seg0
            i[o + 7] = len(rids)
            for k, rid in enumerate(rids):
                i[o + 8 + k] = rid
        ang = 2.0 * math.pi * np.arange(NUM_LASERS, dtype=np.float64) / NUM_LASERS
        ray = np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)
        f[off_ray:off_ray + NUM_LASERS * 2] = ray.reshape(-1)
        return MapTables(name, blob)


def _wrap(h):
    while h > math.pi:
        h -= 2.0 * math.pi
    while h <= -math.pi:
        h += 2.0 * math.pi
    return h


class MapTables:
    """Typed numpy views over a packed blob (what the oracle reads; the kernel reads the blob)."""

    def __init__(self, name, blob):
        self.name = name
        self.blob = np.ascontiguousarray(blob, dtype=np.uint32)
        h = self.blob
        assert int(h[0]) == MAGIC
        self.n_seg, self.n_route, self.n_spawn = int(h[1]), int(h[2]), int(h[3])
        o_seg, o_route, o_spawn, o_ray = int(h[4]), int(h[5]), int(h[6]), int(h[7])
        self.n_ray = int(h[10])
        self.base_obs_dim = int(h[11])
        self.n_side = int(h[12])
        f = self.blob.view(np.float32)
        i = self.blob.view(np.int32)
        self.seg = f[o_seg:o_seg + self.n_seg * SEG_WORDS].reshape(self.n_seg, SEG_WORDS)
        rt_i = i[o_route:o_route + self.n_route * ROUTE_WORDS].reshape(self.n_route, ROUTE_WORDS)
        rt_f = f[o_route:o_route + self.n_route * ROUTE_WORDS].reshape(self.n_route, ROUTE_WORDS)
        self.route_nseg = rt_i[:, 0]
        self.route_seg = rt_i[:, 1:1 + ROUTE_MAX_SEGS]
        self.route_len = rt_f[:, 7]
        sp_f = f[o_spawn:o_spawn + self.n_spawn * SPAWN_WORDS].reshape(self.n_spawn, SPAWN_WORDS)
        sp_i = i[o_spawn:o_spawn + self.n_spawn * SPAWN_WORDS].reshape(self.n_spawn, SPAWN_WORDS)
        self.spawn_f = sp_f[:, 0:6]       # x y h cos sin s0
        self.spawn_seg = sp_i[:, 6]
        self.spawn_nroute = sp_i[:, 7]
        self.spawn_route = sp_i[:, 8:8 + SPAWN_MAX_ROUTES]
        self.ray = f[o_ray:o_ray + self.n_ray * 2].reshape(self.n_ray, 2)

    @property
    def obs_dim(self):
        return self.base_obs_dim

    def bounding_box(self):
        """(x_min, x_max, y_min, y_max) of the drivable surface - what the reference reads from
        `road_network.get_bounding_box()` for the traffic-light message (env_wrappers.py:268-272).  Every segment's
        centre line is sampled at 17 points and widened by its left / right widths."""
        xs, ys = [], []
        for x0, y0, h0, slen, kappa, wl, wr, c0, s0, cx, cy, r in self.seg.astype(np.float64):
            for t in np.linspace(0.0, slen, 17):
                if kappa == 0.0:
                    x, y, h = x0 + t * c0, y0 + t * s0, h0
                else:
                    h = h0 + kappa * t
                    sgn = 1.0 if kappa > 0.0 else -1.0
                    x, y = cx + sgn * r * math.sin(h), cy - sgn * r * math.cos(h)
                nx, ny = -math.sin(h), math.cos(h)                     # left normal
                xs += [x + wl * nx, x - wr * nx]
                ys += [y + wl * ny, y - wr * ny]
        return float(min(xs)), float(max(xs)), float(min(ys)), float(max(ys))


# ---------------------------------------------------------------------------------------------
# Map builders.  Right-hand traffic; arm k points outward along angle k*90 deg.
# ---------------------------------------------------------------------------------------------
def _arm_frames(k):
    a = k * math.pi / 2.0
    u = (math.cos(a), math.sin(a))
    v = (-math.sin(a), math.cos(a))
    return a, u, v


def _lane_bounds(lane, n_lanes=2, w=LANE_WIDTH):
    """Lateral drivable bounds (left, right) of lane `lane` (0 = next to the road centre line)."""
    return (lane + 0.5) * w, (n_lanes - lane - 0.5) * w


def build_intersection(arm_len=60.0, junction=10.0, spawn_spacing=10.0, spawn_first=5.0, spawns_per_lane=5):
    """4-way crossing, two lanes per direction (lane width 3.5, arm 60 m, junction 10 m)."""
    b = _Builder()
    w = LANE_WIDTH
    J = junction
    for arm in range(4):
        a, u, v = _arm_frames(arm)
        for lane in range(2):
            off = (lane + 0.5) * w
            wl, wr = _lane_bounds(lane)
            start = ((J + arm_len) * u[0] + off * v[0], (J + arm_len) * u[1] + off * v[1], a + math.pi)
            routes = []
            turns = (
                ("a", J - off, -math.pi / 2.0),   # right turn, to arm+1
                ("s", 2.0 * J, 0.0),              # straight, to arm+2
                ("a", J + off, +math.pi / 2.0),   # left turn, to arm+3
            )
            for turn in turns:
                if turn[0] == "s":
                    mid = ("s", turn[1], w, w)
                else:
                    mid = ("a", turn[1], turn[2], w, w)
                ids = b.chain(start, [("s", arm_len, wl, wr), mid, ("s", arm_len, wl, wr)])
                routes.append(b.add_route(ids))
            seg0 = b.routes[routes[0]][0][0]
            for q in range(spawns_per_lane):
                b.add_spawn(seg0, spawn_first + q * spawn_spacing, routes)
    return b.pack("intersection")


def build_roundabout(arm_len=60.0, ring_outer=22.0, entry_radius=10.0, spawn_spacing=10.0, spawn_first=5.0,
                     spawns_per_lane=5):
    """4-arm roundabout, two circulating lanes (counter-clockwise)."""
    b = _Builder()
    w = LANE_WIDTH
    re = entry_radius
    for arm in range(4):
        a, u, v = _arm_frames(arm)
        for lane in range(2):
            off = (lane + 0.5) * w
            wl, wr = _lane_bounds(lane)
            rho = ring_outer - (1 - lane) * w          # lane 1 (outer) rides the outer ring lane
            t0 = math.sqrt((rho + re) ** 2 - (off + re) ** 2)
            theta_t = math.atan2(off + re, t0)
            far = t0 + arm_len
            start = (far * u[0] + off * v[0], far * u[1] + off * v[1], a + math.pi)
            routes = []
            for k_exit in (1, 2, 3):
                beta = k_exit * math.pi / 2.0
                ring_ang = beta - 2.0 * theta_t
                ids = b.chain(start, [
                    ("s", arm_len, wl, wr),
                    ("a", re, -(math.pi / 2.0 - theta_t), w, w),
                    ("a", rho, ring_ang, w, w),
                    ("a", re, -(math.pi / 2.0 - theta_t), w, w),
                    ("s", arm_len, wl, wr),
                ])
                routes.append(b.add_route(ids))
            seg0 = b.routes[routes[0]][0][0]
            for q in range(spawns_per_lane):
                b.add_spawn(seg0, spawn_first + q * spawn_spacing, routes)
    return b.pack("roundabout")


def build_tollgate(approach=70.0, gate_len=20.0, leave=70.0, n_lanes=4, spawn_spacing=12.0, spawn_first=5.0,
                   spawns_per_lane=5):
    """Two opposing carriageways of `n_lanes` gated lanes each; every lane is its own route."""
    b = _Builder()
    w = LANE_WIDTH
    for direction in range(2):
        h = 0.0 if direction == 0 else math.pi
        sgn = 1.0 if direction == 0 else -1.0
        x_start = -sgn * (approach + gate_len / 2.0)
        for lane in range(n_lanes):
            off = (lane + 0.5) * w
            # right-hand traffic: lanes sit on the right of the centre line of travel
            y = -sgn * off
            wl, wr = _lane_bounds(lane, n_lanes)
            ids = b.chain((x_start, y, h), [("s", approach, wl, wr), ("s", gate_len, 0.5 * w, 0.5 * w),
                                            ("s", leave, wl, wr)])
            rid = b.add_route(ids)
            for q in range(spawns_per_lane):
                b.add_spawn(ids[0], spawn_first + q * spawn_spacing, [rid])
    return b.pack("tollgate", base_obs_dim=156, n_side=65)


def build_bottleneck(wide=60.0, taper=25.0, narrow=30.0, n_wide=4, spawn_spacing=11.0, spawn_first=5.0,
                     spawns_per_lane=3):
    """`n_wide` lanes funnel into one lane per direction and widen again."""
    b = _Builder()
    w = LANE_WIDTH
    for direction in range(2):
        h = 0.0 if direction == 0 else math.pi
        sgn = 1.0 if direction == 0 else -1.0
        x_start = -sgn * (wide + taper + narrow / 2.0)
        for lane in range(n_wide):
            off = (lane + 0.5) * w
            wl, wr = _lane_bounds(lane, n_wide)
            shift = off - 0.5 * w              # lateral move needed to reach the single narrow lane
            ang = math.atan2(shift, taper)
            hyp = math.hypot(shift, taper)
            pieces = [("s", wide, wl, wr)]
            # funnel: a straight diagonal towards the narrow lane, then the narrow lane, then fan out again
            start = (x_start, -sgn * off, h)
            ids = b.chain(start, pieces)
            x1, y1, _ = _Builder._end_pose(start[0], start[1], h, wide, 0.0)
            ids += b.chain((x1, y1, h + ang), [("s", hyp, 1.5 * w, 1.5 * w)])
            x2 = x1 + sgn * taper
            y2 = -sgn * 0.5 * w
            ids += b.chain((x2, y2, h), [("s", narrow, 0.5 * w, 0.5 * w)])
            x3 = x2 + sgn * narrow
            ids += b.chain((x3, y2, h - ang), [("s", hyp, 1.5 * w, 1.5 * w)])
            x4 = x3 + sgn * taper
            ids += b.chain((x4, -sgn * off, h), [("s", wide, wl, wr)])
            rid = b.add_route(ids)
            for q in range(spawns_per_lane):
                b.add_spawn(ids[0], spawn_first + q * spawn_spacing, [rid])
    return b.pack("bottleneck", base_obs_dim=96, n_side=5)


def build_parking_lot(parking_space_num=8, road_len=60.0, slot_depth=8.0, spawn_first=6.0):
    """One two-way road with `parking_space_num` perpendicular bays; cars drive road->bay or bay->road."""
    b = _Builder()
    w = LANE_WIDTH
    r_turn = 5.0
    per_side = parking_space_num // 2
    pitch = 2.0 * r_turn + 2.0
    x_first = -(per_side - 1) * pitch / 2.0
    for direction in range(2):
        h = 0.0 if direction == 0 else math.pi
        sgn = 1.0 if direction == 0 else -1.0
        y = -sgn * 0.5 * w
        x_start = -sgn * road_len
        road_routes = []
        # through route
        ids = b.chain((x_start, y, h), [("s", 2.0 * road_len, 0.5 * w, 0.5 * w)])
        road_routes.append(b.add_route(ids))
        # road -> bay on the right-hand side: straight, right turn, straight into the bay
        for k in range(per_side):
            xb = x_first + k * pitch
            run = (xb - r_turn * sgn) - x_start
            run = run * sgn
            if run <= 1.0 or len(road_routes) >= SPAWN_MAX_ROUTES:
                continue
            ids = b.chain((x_start, y, h), [("s", run, 0.5 * w, 0.5 * w), ("a", r_turn, -math.pi / 2.0, w, w),
                                            ("s", slot_depth, 0.5 * w, 0.5 * w)])
            road_routes.append(b.add_route(ids))
        b.add_spawn(b.routes[road_routes[0]][0][0], spawn_first, road_routes[:1])
        # bay -> road: leave the bay forwards, right turn onto the near lane, drive off
        for k in range(per_side):
            xb = x_first + k * pitch
            yb = -sgn * (0.5 * w + r_turn + slot_depth)
            hb = h + math.pi / 2.0          # facing the road
            # near lane after a right turn travels in direction (h + pi/2 - pi/2) = h
            leave = road_len - sgn * (xb + sgn * r_turn)
            ids = b.chain((xb, yb, hb), [("s", slot_depth, 0.5 * w, 0.5 * w), ("a", r_turn, -math.pi / 2.0, w, w),
                                         ("s", leave, 0.5 * w, 0.5 * w)])
            rid = b.add_route(ids)
            b.add_spawn(ids[0], 1.0, [rid])
    # extra road spawn places so that num_agents=10 always fits
    for direction in range(2):
        seg0 = b.spawns[0][4] if direction == 0 else b.spawns[1 + per_side][4]
        rts = b.spawns[0][5] if direction == 0 else b.spawns[1 + per_side][5]
        for q in range(1, 4):
            b.add_spawn(seg0, spawn_first + 10.0 * q, rts)
    return b.pack("parking_lot")


def build_pg(seed=0, num_blocks=3, straight=50.0, radius=35.0, spawn_spacing=10.0, spawn_first=5.0,
             spawns_per_lane=4):
    """Procedurally generated two-way road (the reference's `MultiAgentMetaDrive` runs on MetaDrive's PG maps): a
    straight entry, `num_blocks` blocks drawn from the map seed - straight, left curve, right curve of 30 to 90
    degrees - and a straight exit; two lanes per direction, every lane one route, spawn places on the entry straight
    of each direction."""
    assert 0 <= num_blocks <= ROUTE_MAX_SEGS - 2, "a route holds at most %d segments" % ROUTE_MAX_SEGS
    rng = np.random.default_rng(int(seed))
    blocks = []
    for _ in range(num_blocks):
        kind = int(rng.integers(0, 3))
        if kind == 0:
            blocks.append(("s", straight))
        else:
            ang = float(rng.uniform(math.pi / 6.0, math.pi / 2.0)) * (1.0 if kind == 1 else -1.0)
            blocks.append(("a", radius, ang))
    centre = [("s", straight)] + blocks + [("s", straight)]
    # end pose of the centre line (the reverse carriageway starts there, facing back)
    x, y, h = 0.0, 0.0, 0.0
    for p in centre:
        if p[0] == "s":
            x, y, h = _Builder._end_pose(x, y, h, p[1], 0.0)
        else:
            x, y, h = _Builder._end_pose(x, y, h, abs(p[2]) * p[1], (1.0 if p[2] > 0 else -1.0) / p[1])
    b = _Builder()
    w = LANE_WIDTH
    for direction, (sx, sy, sh, pieces) in enumerate(((0.0, 0.0, 0.0, centre),
                                                      (x, y, h + math.pi, [q if q[0] == "s" else ("a", q[1], -q[2])
                                                                           for q in reversed(centre)]))):
        for lane in range(2):
            off = (lane + 0.5) * w                     # to the right of the centre line in the travel direction
            wl, wr = _lane_bounds(lane)
            start = (sx + off * math.sin(sh), sy - off * math.cos(sh), sh)
            lane_pieces = []
            for q in pieces:
                if q[0] == "s":
                    lane_pieces.append(("s", q[1], wl, wr))
                else:                                  # a lane right of the centre line: wider on left curves
                    lane_pieces.append(("a", q[1] + off if q[2] > 0 else q[1] - off, q[2], wl, wr))
            ids = b.chain(start, lane_pieces)
            rid = b.add_route(ids)
            for k in range(spawns_per_lane):
                b.add_spawn(ids[0], spawn_first + k * spawn_spacing, [rid])
    return b.pack("pg")


_BUILDERS = {
    "pg": build_pg,
    "intersection": build_intersection,
    "roundabout": build_roundabout,
    "tollgate": build_tollgate,
    "bottleneck": build_bottleneck,
    "parking_lot": build_parking_lot,
}
_CACHE = {}


def build_map(name, **kwargs):
    key = (name, tuple(sorted(kwargs.items())))
    if key not in _CACHE:
        _CACHE[key] = _BUILDERS[name](**kwargs)
    return _CACHE[key]
