"""ref_transcript_rods.py — the independent restatement (see ref_transcript.py) extended to the bacterial-rods agent.

TEST INFRASTRUCTURE.  Written directly from the Rust sources of the reference WITHOUT reading ``oracle/``:

* ``nearest_point_from_point_to_line``   cellular_raza-building-blocks/src/cell_building_blocks/interaction.rs:497-509
* ``RodInteraction::calculate_force_between``   …/bacterial_rods.rs:292-336 (around ``MorsePotential``, interaction.rs:151-182)
* ``RodMechanics::calculate_increment`` / ``get_random_contribution``   …/bacterial_rods.rs:157-240
* ``RodMechanics::divide``   …/bacterial_rods.rs:586-663
* ``CartesianSubDomainRods::apply_boundary`` / ``get_voxel_index_of``   …/bacterial_rods.rs:489-555
* the example agent (NeighborSensing, Cycle)   cellular_raza-examples/bacterial_rods/src/main.rs:48-109
* where chili calls the neighbour sensing   cellular_raza-core/src/backend/chili/update_mechanics.rs:65-108, 153-238

The chili step itself (pair weights, queues, Adams–Bashforth wrapper, cell sorting) is ``ChiliTranscript`` of
ref_transcript.py, which is generic over the agent: a rod is a flat list of 3 V floats (vertex-major: v * 3 + c)
wherever chili only adds, scales or moves it, and a V x 3 matrix inside the rod-specific functions below.
Divisions are transcribed as a function (``divide``) and checked on their own; the stepping transcript refuses a
step in which a rod would divide.
"""
from __future__ import annotations

import math

import ref_transcript as R


def rows(flat, V):
    return [list(flat[v * 3:(v + 1) * 3]) for v in range(V)]


def flat(mat):
    return [x for row in mat for x in row]


def dot(a, b):
    s = a[0] * b[0]
    for k in range(1, len(a)):
        s = s + a[k] * b[k]
    return s


def norm(a):
    return math.sqrt(dot(a, a))


# ---- interaction.rs:497-509 -----------------------------------------------------------------------------------
def nearest_point_from_point_to_line(point, line):
    a, b = line
    ab = R.vsub(b, a)
    ap = R.vsub(point, a)
    t = dot(ab, ap) / dot(ab, ab)
    t = 0.0 if t < 0.0 else (1.0 if t > 1.0 else t)  # f64::clamp (a NaN stays a NaN)
    nearest = R.vadd(R.vscale(a, 1.0 - t), R.vscale(b, t))
    return norm(R.vsub(point, nearest)), nearest, t


# ---- bacterial_rods.rs:292-336 ---------------------------------------------------------------------------------
class RodInteraction:
    def __init__(self, inner, V):
        self.inner = inner  # the point interaction (MorsePotential of ref_transcript.py)
        self.V = V

    def get_interaction_information(self):
        return self.inner.get_interaction_information()

    def calculate_force_between(self, own_pos, own_vel, ext_pos, ext_vel, ext_inf):
        V = self.V
        op, ov, ep, ev = rows(own_pos, V), rows(own_vel, V), rows(ext_pos, V), rows(ext_vel, V)
        force_own = [[0.0] * 3 for _ in range(V)]
        force_ext = [[0.0] * 3 for _ in range(V)]
        for i, p1 in enumerate(op):
            for j in range(V - 1):  # tuple_windows over the rows of ext_pos
                p2_n0, p2_n1 = ep[j], ep[j + 1]
                _, nearest_point, rel_length = nearest_point_from_point_to_line(p1, (p2_n0, p2_n1))
                f_own, f_ext = self.inner.calculate_force_between(p1, ov[i], nearest_point, ev[j], ext_inf)
                force_own[i] = R.vadd(force_own[i], f_own)
                force_ext[j] = R.vadd(force_ext[j], R.vscale(f_ext, 1.0 - rel_length))
                k = (j + 1) % V
                force_ext[k] = R.vadd(force_ext[k], R.vscale(f_ext, rel_length))
        return flat(force_own), flat(force_ext)


# ---- bacterial_rods.rs:82-107, 157-240 ---------------------------------------------------------------------------
class RodMechanics:
    def __init__(self, pos, vel, diffusion_constant, spring_tension, rigidity, spring_length, damping, V):
        self.pos, self.vel = list(pos), list(vel)
        self.diffusion_constant, self.spring_tension, self.rigidity = diffusion_constant, spring_tension, rigidity
        self.spring_length, self.damping, self.V = spring_length, damping, V

    def velocity(self):
        return list(self.vel)

    def set_velocity(self, v):
        self.vel = list(v)

    def calculate_increment(self, force):
        one_half = 1.0 / (1.0 + 1.0)
        two = 1.0 + 1.0
        V = self.V
        total = rows(force, V)
        pos = rows(self.pos, V)
        dist_internal = [R.vsub(pos[i], pos[i + 1]) for i in range(V - 1)]
        for i, dist in enumerate(dist_internal):
            if norm(dist) != 0.0:
                n = norm(dist)
                dirv = [x / n for x in dist]  # normalize() = self / self.norm()
                force_internal = R.vscale(R.vscale(R.vneg(dirv), self.spring_tension), norm(dist) - self.spring_length)
                total[i] = R.vadd(total[i], R.vscale(force_internal, one_half))
                total[i + 1] = R.vadd(total[i + 1], R.vscale(R.vneg(force_internal), one_half))
        for i in range(V - 2):  # tuple_windows over dist_internal
            conn1, conn2 = dist_internal[i], dist_internal[i + 1]
            # nalgebra Matrix::angle: 0 if a norm is zero, else acos(clamp(dot / (n1 * n2), -1, 1))
            n1, n2 = norm(conn1), norm(conn2)
            if n1 == 0.0 or n2 == 0.0:
                angle = 0.0
            else:
                cang = dot(conn1, conn2) / (n1 * n2)
                cang = -1.0 if cang < -1.0 else (1.0 if cang > 1.0 else cang)
                angle = math.acos(cang)
            force_d = R.vsub(conn1, conn2)
            nd = norm(force_d)
            force_dir = [x / nd for x in force_d] if nd != 0.0 else force_d
            f = R.vscale(R.vscale(R.vscale(force_dir, two), self.rigidity), math.tan(one_half * angle))
            total[i] = R.vadd(total[i], R.vscale(R.vneg(f), one_half))
            total[i + 1] = R.vadd(total[i + 1], f)
            total[i + 2] = R.vadd(total[i + 2], R.vscale(R.vneg(f), one_half))
        vel = rows(self.vel, V)
        total = [R.vsub(t, R.vscale(v, self.damping)) for t, v in zip(total, vel)]
        return list(self.vel), flat(total)

    def get_random_contribution(self, z_row, dt):
        """dpos = from_distribution(nrows, Normal(0, sqrt(dt))) * 2^-2 * D / dt; nalgebra fills the matrix column by
        column, so the k-th number drawn goes to (row k % V, column k // V); a Normal sample is mean + std * z"""
        V = self.V
        sd = math.sqrt(dt)
        m = [[0.0] * 3 for _ in range(V)]
        for k in range(3 * V):
            m[k % V][k // V] = 0.0 + sd * z_row[k]
        quarter = (1.0 + 1.0) ** -2
        dpos = [[((x * quarter) * self.diffusion_constant) / dt for x in row] for row in m]
        return flat(dpos), [0.0] * (3 * V)

    # ---- bacterial_rods.rs:586-663 -----------------------------------------------------------------------------
    def divide(self, radius):
        """-> (positions of c1 = self after the division, positions of the new cell c2, new spring length)"""
        V = self.V
        pos = rows(self.pos, V)
        two = 1.0 + 1.0
        one_half = 1.0 / two
        div_factor = one_half - radius / (float(V) * self.spring_length)
        new_spring_length = div_factor * self.spring_length
        c1, c2 = [list(r) for r in pos], [list(r) for r in pos]
        segments = [norm(R.vsub(pos[i], pos[i + 1])) for i in range(V - 1)]

        def fsum(xs):  # Iterator::sum::<f64>()
            s = 0.0
            for x in xs:
                s = s + x
            return s

        segment_length = (fsum(segments) - two * radius) / float(V - 1) / two
        rev = segments[::-1]
        for n_vertex in range(V):
            target = float(n_vertex) * segment_length
            ks = [n for n in range(len(segments)) if fsum(segments[:n]) <= target]
            if not ks:
                raise R.CalcError("DivisionError: Could not calculate new segment lengths")
            k = max(ks)
            q = (target - fsum(segments[:k])) / segments[k]
            c1[n_vertex] = R.vadd(R.vscale(pos[k], 1.0 - q), R.vscale(pos[k + 1], q))
            ms = [n for n in range(V) if fsum(rev[:n]) <= target]
            if not ms:
                raise R.CalcError("DivisionError: Could not calculate new segment lengths")
            m = max(ms)
            if V - m - 2 < 0:
                raise R.CalcError("DivisionError: index underflow (a panic in the reference)")
            p = (target - fsum(rev[:m])) / segments[V - m - 2]
            c2[V - n_vertex - 1] = R.vadd(R.vscale(pos[V - m - 1], 1.0 - p), R.vscale(pos[V - m - 2], p))
        return c1, c2, new_spring_length


# ---- the example agent: main.rs:12-109 ----------------------------------------------------------------------------
class RodAgentInteraction(RodInteraction):
    pass


class RodCellState:
    """growth_rate, spring_length_threshold and GROWTH_BASE_RATE of the example's Agent (main.rs:12-25)"""

    def __init__(self, growth_rate, spring_length_threshold, growth_base_rate):
        self.growth_rate, self.threshold, self.base = growth_rate, spring_length_threshold, growth_base_rate


def accumulate_information(own_pos, ext_pos, radius, V):
    acc = 0
    for own_point in rows(own_pos, V):
        for ext_point in rows(ext_pos, V):
            if norm(R.vsub(own_point, ext_point)) < 2.0 * radius:
                acc += 1
    return acc


def react_to_neighbors(state: RodCellState, neighbors: int):
    if neighbors > 0:
        state.growth_rate = max(state.base * (8.0 - float(neighbors)) / 8.0, 0.0)
    else:
        state.growth_rate = state.base


# ---- CartesianCuboidRods: bacterial_rods.rs:368-372, 489-555 -------------------------------------------------------
class RodDomain:
    """what ChiliTranscript asks of a domain, for rods: the voxel of a rod is the voxel of the mean of its vertices
    (row_sum / nrows), the boundary acts on every vertex"""

    def __init__(self, cuboid: R.CartesianCuboid, V: int):
        self.c, self.V = cuboid, V
        self.D = 3 * V  # length of the flat position / force vectors chili moves around

    def create_subdomains(self, n):
        return self.c.create_subdomains(n)

    def get_neighbor_voxel_indices(self, v):
        return self.c.get_neighbor_voxel_indices(v)

    def _mean(self, pos):
        r = rows(pos, self.V)
        s = [0.0, 0.0, 0.0]
        for row in r:  # row_sum: fold(0, +) down every column
            s = R.vadd(s, row)
        return [x / float(self.V) for x in s]

    def get_voxel_index_of_raw(self, pos):
        return self.c.get_voxel_index_of_raw(self._mean(pos))

    def get_index_of(self, pos):
        return self.c.get_index_of(self._mean(pos))

    def apply_boundary(self, pos, vel):
        p, v = rows(pos, self.V), rows(vel, self.V)
        two = 1.0 + 1.0
        for pr, vr in zip(p, v):
            for i in range(3):
                if pr[i] < self.c.min[i]:
                    pr[i] = self.c.min[i] * two - pr[i]
                    vr[i] = abs(vr[i])
                if pr[i] > self.c.max[i]:
                    pr[i] = self.c.max[i] * two - pr[i]
                    vr[i] = -abs(vr[i])
        for pr in p:
            for i in range(3):
                if pr[i] < self.c.min[i] or pr[i] > self.c.max[i]:
                    raise R.BoundaryError("Particle is out of domain")
        return flat(p), flat(v)


class RodsTranscript(R.ChiliTranscript):
    """aspects [Mechanics, Interaction, Cycle, NeighborSensing] of the example (main.rs:202-210) without Division events:
    forces and neighbour counts by the generic chili steps, then per cell local_mechanics_update,
    local_interaction_react_to_neighbors, local_cycle_update (run_sim.rs:262-283), then boundary and sorting."""

    def __init__(self, cuboid, V, n_threads, dt, solver_order=2):
        super().__init__(RodDomain(cuboid, V), n_threads, dt, solver_order)
        self.V = V

    def neighbor_counts(self):
        """update_mechanics.rs:65-108, 186-217: inside a voxel both cells of a pair sense each other; across voxels the
        cells of the RECEIVING voxel sense the sender — every cell of every neighbouring voxel sends once, so a cell
        ends up with the sum over all other cells of its 3^D neighbourhood"""
        voxels = {}
        for box in self.boxes:
            voxels.update(box.voxels)
        counts = {}
        for p, vox in voxels.items():
            for cell, _ in vox.cells:
                acc = 0
                for q in [p] + list(vox.neighbors):
                    for other, _ in voxels[q].cells:
                        if other is cell:
                            continue
                        acc += accumulate_information(cell.mech.pos, other.mech.pos, cell.inter.inner.radius, self.V)
                counts[cell.key] = acc
        return counts

    def step(self, noise_row=None, states=None):
        counts = self.neighbor_counts()
        self.force_steps()
        self.update_steps(noise_row)  # local_mechanics_update for every cell, boundary, sorting
        # the two other local functions only touch per-cell state that the mechanics update does not read
        for _, cell, _ in self.cells_in_order():
            st = states[cell.key]
            react_to_neighbors(st, counts[cell.key])
            cell.mech.spring_length += st.growth_rate * self.dt
            if cell.mech.spring_length > st.threshold:
                raise NotImplementedError("CycleEvent::Division inside the stepping transcript")
