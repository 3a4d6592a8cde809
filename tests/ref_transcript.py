"""ref_transcript.py — a SECOND, independent restatement of the chili mechanics/interaction step.

TEST INFRASTRUCTURE.  Written directly from the Rust sources of the reference (file:line cited at every
function, paths relative to /root/reference) WITHOUT reading ``oracle/``: the C++ oracle and the CUDA
kernels were written by one hand from the same pseudocode, so a shared misreading would be invisible to
a device-vs-oracle test.  This transcript is deliberately literal and slow: Python objects per cell,
one ``Voxel`` per plain index, one ``SubDomainBox`` per emulated thread with explicit message queues
(``PosInformation`` / ``ForceInformation`` / ``SendCell``), two ``calculate_force_between`` calls per
pair, the dropped local return force, the push-before-read ring buffer.  It is used to

* generate ``tests/golden/ref_derived_*.json`` (``tools/make_ref_derived.py``),
* check oracle == transcript on the CPU (``tests/test_ref_transcript.py``),
* check device == fixtures on the GPU (``tests/test_gpu_ref_derived.py``).

Arithmetic: Python floats are IEEE-754 binary64 and every operation below is written in the operation
order of the Rust expression it follows (Rust never contracts ``a * b + c``).  nalgebra (un-vendored,
0.34): ``norm() = sqrt(sum of squares in index order)``, ``normalize() = self / self.norm()``,
``v / s`` and ``v * s`` component-wise.  ``powf`` is libm ``pow`` (Python ``**`` on floats).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

# --------------------------------------------------------------------------------------------------------
# small vector helpers (lists of floats; index order is the summation order)
# --------------------------------------------------------------------------------------------------------
Vec = list


def vsub(a, b):
    return [x - y for x, y in zip(a, b)]


def vadd(a, b):
    return [x + y for x, y in zip(a, b)]


def vscale(a, s):
    return [x * s for x in a]


def vdiv(a, s):
    return [x / s for x in a]


def vneg(a):
    return [-x for x in a]


def vnorm(a):
    s = 0.0
    for x in a:
        s = s + x * x
    return math.sqrt(s)


def vzeros(d):
    return [0.0] * d


def xapy(x, a, y):
    """cellular_raza-concepts/src/reactions.rs:74-76  ``self * a + y``"""
    return [xi * a + yi for xi, yi in zip(x, y)]


def xa(x, a):
    """cellular_raza-concepts/src/reactions.rs:78-80  ``self * a``"""
    return [xi * a for xi in x]


def signum(x: float) -> float:
    """f64::signum: 1.0 for +0.0 and positive, -1.0 for -0.0 and negative, NaN for NaN"""
    if x != x:
        return x
    return math.copysign(1.0, x)


class CalcError(Exception):
    pass


class BoundaryError(Exception):
    pass


class IndexError_(Exception):
    pass


# --------------------------------------------------------------------------------------------------------
# interactions
# --------------------------------------------------------------------------------------------------------
@dataclass
class CellSpecificInteraction:
    """cellular_raza-examples/cell_sorting/src/main.rs:31-103 (``bench=False``) and the copy with the
    opposite sign convention cellular_raza-benchmarks/src/cell_sorting.rs:34-88 (``bench=True``)."""
    species: int
    cell_radius: float
    potential_strength: float
    relative_interaction_range: float
    bench: bool = False

    def get_interaction_information(self):
        return (self.cell_radius, self.species)

    def calculate_force_between(self, own_pos, own_vel, ext_pos, ext_vel, ext_info):
        min_relative_distance_to_center = 0.3162277660168379
        if not (vnorm(vsub(own_pos, ext_pos)) < self.cell_radius * min_relative_distance_to_center):
            z = vsub(own_pos, ext_pos)
            r = vnorm(z)
            dir_ = vdiv(z, vnorm(z))
        else:
            if own_pos == ext_pos:
                return vzeros(len(own_pos)), vzeros(len(own_pos))
            z = vsub(own_pos, ext_pos)
            dir_ = vdiv(z, vnorm(z))
            r = self.cell_radius * min_relative_distance_to_center
        ext_radius, species = ext_info
        sigma = r / (self.cell_radius + ext_radius)
        bound = 4.0 + 1.0 / sigma
        spatial_cutoff = (1.0 + signum(self.relative_interaction_range * (self.cell_radius + ext_radius) - r)) * 0.5
        inner = (1.0 / sigma) ** 2.0 - (1.0 / sigma) ** 4.0
        # f64::min / f64::max (ignore a NaN operand; no NaN can reach here)
        inner = min(inner, bound)
        inner = max(inner, -bound)
        strength = self.potential_strength * inner
        attracting_force = vscale(vscale(dir_, max(strength, 0.0)), spatial_cutoff)
        repelling_force = vscale(vscale(dir_, min(strength, 0.0)), spatial_cutoff)
        if species == self.species:
            a = vsub(vneg(repelling_force), attracting_force)
            b = vadd(repelling_force, attracting_force)
        else:
            a = vneg(repelling_force)
            b = list(repelling_force)
        if self.bench:  # cell_sorting.rs:79-86: (rep + att, -rep - att) / (rep, -rep)
            return b, a
        return a, b


@dataclass
class MorsePotential:
    """cellular_raza-building-blocks/src/cell_building_blocks/interaction.rs:151-182, 226-260"""
    radius: float
    potential_stiffness: float
    cutoff: float
    strength: float

    def get_interaction_information(self):
        return self.radius

    def calculate_force_between(self, own_pos, own_vel, ext_pos, ext_vel, ext_info):
        z = vsub(own_pos, ext_pos)
        dist = vnorm(z)
        if dist > self.cutoff or dist == 0.0:
            return vzeros(len(z)), vzeros(len(z))
        dir_ = vdiv(z, dist)
        r = self.radius + ext_info
        s = self.strength
        a = self.potential_stiffness
        two = 1.0 + 1.0
        e = math.exp(-a * (dist - r))
        force = -two * s * a * e * (1.0 - e)
        return vscale(dir_, force), vscale(vneg(dir_), force)


@dataclass
class BoundLennardJones:
    """interaction.rs:112-145: no guard for r == 0 (NaN), hard cutoff ``cutoff >= r``"""
    epsilon: float
    sigma: float
    bound: float
    cutoff: float

    def get_interaction_information(self):
        return None

    def calculate_force_between(self, own_pos, own_vel, ext_pos, ext_vel, ext_info):
        z = vsub(own_pos, ext_pos)
        r = vnorm(z)
        try:
            dir_ = vdiv(z, r)
            val = 4.0 * self.epsilon / r * (12.0 * (self.sigma / r) ** 11.0 - 6.0 * (self.sigma / r) ** 5.0)
            mx = self.bound / r
        except ZeroDivisionError:  # IEEE: 0/0 = NaN propagates through the reference's expression
            nan = float("nan")
            return [nan] * len(z), [nan] * len(z)
        q = 1.0 if self.cutoff >= r else 0.0
        m = min(mx, val)
        return vscale(vscale(vneg(dir_), q), m), vscale(vscale(dir_, q), m)


@dataclass
class MiePotential:
    """interaction.rs:330-389"""
    radius: float
    strength: float
    bound: float
    cutoff: float
    en: float
    em: float

    def get_interaction_information(self):
        return self.radius

    def radius_to_sigma_factor(self):
        return (self.em / self.en) ** (1.0 / (self.en - self.em))

    def calculate_force_between(self, own_pos, own_vel, ext_pos, ext_vel, ext_radius):
        z = vsub(own_pos, ext_pos)
        r = vnorm(z)
        if r == 0.0:
            raise CalcError("identical position for two objects. Cannot Calculate force in Mie potential")
        if r > self.cutoff:
            return vzeros(len(z)), vzeros(len(z))
        dir_ = vdiv(z, r)
        x = self.radius + ext_radius
        sigma = self.radius_to_sigma_factor() * x
        mie_constant = self.en / (self.en - self.em) * (self.en / self.em) ** (self.em / (self.en - self.em))
        potential_part = self.en / sigma * (sigma / r) ** (self.en + 1.0) - self.em / sigma * (sigma / r) ** (self.em + 1.0)
        force = min(mie_constant * self.strength * potential_part, self.bound)
        return vscale(dir_, force), vscale(vneg(dir_), force)


# --------------------------------------------------------------------------------------------------------
# mechanics (cellular_raza-building-blocks/src/cell_building_blocks/mechanics.rs)
# --------------------------------------------------------------------------------------------------------
def wiener_process(z_row: Sequence[float], dt: float):
    """mechanics.rs:211-226 with the ChaCha8/ziggurat draw replaced by the shared noise table:
    ``Normal::new(0, sqrt(dt))`` samples ``mean + std_dev * z`` (rand_distr 0.5), then ``random_dir / dt``."""
    std_dev = math.sqrt(dt)
    random_dir = [0.0 + std_dev * z for z in z_row]
    return vdiv(random_dir, dt)


@dataclass
class NewtonDamped:
    """mechanics.rs:100-204"""
    pos: Vec
    vel: Vec
    damping_constant: float
    mass: float

    def velocity(self):
        return list(self.vel)

    def set_velocity(self, v):
        self.vel = list(v)

    def get_random_contribution(self, z_row, dt):
        return vzeros(len(self.pos)), vzeros(len(self.pos))

    def calculate_increment(self, force):
        dx = list(self.vel)
        dv = vsub(vdiv(force, self.mass), vscale(self.vel, self.damping_constant))
        # ``self.damping_constant * self.vel``: scalar * vector, component-wise product (commutative)
        return dx, dv


@dataclass
class Brownian:
    """mechanics.rs:270-392: velocity() is zero, set_velocity is a no-op"""
    pos: Vec
    diffusion_constant: float
    kb_temperature: float

    def velocity(self):
        return vzeros(len(self.pos))

    def set_velocity(self, v):
        pass

    def get_random_contribution(self, z_row, dt):
        w = wiener_process(z_row, dt)
        s = math.sqrt(2.0 * self.diffusion_constant)
        return [s * x for x in w], vzeros(len(self.pos))

    def calculate_increment(self, force):
        f = self.diffusion_constant / self.kb_temperature
        return [f * x for x in force], vzeros(len(self.pos))


@dataclass
class Langevin:
    """mechanics.rs:395-500"""
    pos: Vec
    vel: Vec
    mass: float
    damping: float
    kb_temperature: float

    def velocity(self):
        return list(self.vel)

    def set_velocity(self, v):
        self.vel = list(v)

    def get_random_contribution(self, z_row, dt):
        w = wiener_process(z_row, dt)
        s = math.sqrt(2.0 * self.damping * self.kb_temperature / self.mass)
        return vzeros(len(self.pos)), [s * x for x in w]

    def calculate_increment(self, force):
        dx = list(self.vel)
        dv1 = [1.0 / self.mass * x for x in force]
        dv2 = [-self.damping * x for x in self.vel]
        return dx, vadd(dv1, dv2)


# --------------------------------------------------------------------------------------------------------
# AuxStorageMechanics (cellular_raza-core/src/backend/chili/aux_storage.rs:41-130); RingBuffer = circ_buffer
# 0.1.8 (un-vendored): push appends and evicts the OLDEST when full, iteration is oldest -> newest (pinned by the
# doctest aux_storage.rs:862-868, 894-903)
# --------------------------------------------------------------------------------------------------------
class RingBuffer:
    def __init__(self, n):
        self.n = n
        self.items = []

    def push(self, x):
        if self.n == 0:
            return
        if len(self.items) == self.n:
            self.items.pop(0)
        self.items.append(x)

    def get_size(self):
        return len(self.items)

    def iter(self):
        return iter(self.items)


class AuxStorageMechanics:
    def __init__(self, dim, n):
        self.positions = RingBuffer(n)
        self.velocities = RingBuffer(n)
        self.current_force = vzeros(dim)
        self.zero_force = vzeros(dim)
        # test instrumentation (not part of the reference): sum of |contribution|_inf, the guard term of the
        # force metric of SURVEY.md 8d for agents whose net force cancels
        self.abs_sum = 0.0

    def add_force(self, f):
        self.current_force = vadd(self.current_force, f)
        self.abs_sum += max(abs(x) for x in f) if f else 0.0

    def get_current_force_and_reset(self):
        f = list(self.current_force)
        self.current_force = list(self.zero_force)
        return f


# --------------------------------------------------------------------------------------------------------
# solvers (cellular_raza-core/src/backend/chili/solvers.rs)
# --------------------------------------------------------------------------------------------------------
def euler(x, dx, dt, dx_rand):
    """solvers.rs:303-311"""
    return xapy(xapy(dx, dt, x), 1.0, xa(dx_rand, dt))


def adams_bashforth_3(x, dx, dt, dx_rand):
    """solvers.rs:313-331"""
    f0 = 23.0 / 12.0
    f1 = -16.0 / 12.0
    f2 = 5.0 / 12.0
    dx0, dx1, dx2 = dx
    return xapy(xapy(xapy(dx0, f0, xapy(dx1, f1, xa(dx2, f2))), dt, x), 1.0, xa(dx_rand, dt))


def adams_bashforth_2(x, dx, dt, dx_rand):
    """solvers.rs:333-347"""
    f0 = 3.0 / 2.0
    f1 = -1.0 / 2.0
    dx0, dx1 = dx
    return xapy(xapy(xapy(dx0, f0, xa(dx1, f1)), dt, x), 1.0, xa(dx_rand, dt))


def mechanics_update(order, cell, aux, dt, z_row):
    """MechanicsAdamsBashforthSolver<N>::update (solvers.rs:80-139): N = 2 -> mechanics_adams_bashforth_3
    (:156-231), N = 1 -> mechanics_adams_bashforth_2 (:247-300), N = 0 -> mechanics_euler (:23-55)."""
    force = aux.get_current_force_and_reset()
    velocity = cell.mech.velocity()
    position = list(cell.mech.pos)
    dx, dv = cell.mech.calculate_increment(force)
    dx_rand, dv_rand = cell.mech.get_random_contribution(z_row, dt)
    # "Update values in the aux_storage": the push happens BEFORE the history is read
    aux.positions.push(list(dx))
    aux.velocities.push(list(dv))
    if order == 0:
        new_position = euler(position, dx, dt, dx_rand)
        new_velocity = euler(velocity, dv, dt, dv_rand)
    else:
        n_previous_values = aux.positions.get_size()
        old_pos = aux.positions.iter()
        old_vel = aux.velocities.iter()
        if order == 2 and n_previous_values == 2:
            new_position = adams_bashforth_3(position, [dx, next(old_pos), next(old_pos)], dt, dx_rand)
            new_velocity = adams_bashforth_3(velocity, [dv, next(old_vel), next(old_vel)], dt, dv_rand)
        elif n_previous_values == 1:
            new_position = adams_bashforth_2(position, [dx, next(old_pos)], dt, dx_rand)
            new_velocity = adams_bashforth_2(velocity, [dv, next(old_vel)], dt, dv_rand)
        else:
            new_position = euler(position, dx, dt, dx_rand)
            new_velocity = euler(velocity, dv, dt, dv_rand)
    cell.mech.pos = list(new_position)
    cell.mech.set_velocity(new_velocity)


# --------------------------------------------------------------------------------------------------------
# domain (cellular_raza-building-blocks/src/domains/cartesian_cuboid_n.rs)
# --------------------------------------------------------------------------------------------------------
def get_decomp_res(n_voxel: int, n_regions: int):
    """cartesian_cuboid_n.rs:19-59"""
    average_len = math.ceil(n_voxel / n_regions)

    def residue(n, m, avg):
        return n_voxel - avg * n - (avg - 1) * m

    n = n_regions
    m = 0
    for _ in range(n_regions):
        r = residue(n, m, average_len)
        if r == 0:
            return n, m, average_len
        elif r > 0:
            if n == n_regions:
                average_len += 1
                n = n_regions
                m = 0
        else:
            n -= 1
            m += 1
    return None


def to_usize(x: float) -> Optional[int]:
    """num-traits 0.2 ``ToPrimitive::to_usize`` for f64: Some(trunc) iff -1 < x < 2^64, None for NaN"""
    if x != x:
        return None
    if x > -1.0 and x < 18446744073709551616.0:
        return int(x) if x > 0.0 else 0
    return None


class CartesianCuboid:
    def __init__(self, mn, mx, n_voxels, dx):
        self.min, self.max, self.n_voxels, self.dx = list(mn), list(mx), list(n_voxels), list(dx)
        self.D = len(mn)

    @staticmethod
    def from_boundaries_and_interaction_range(mn, mx, interaction_range):
        """:183-222"""
        n_voxels, dx = [], []
        for i in range(len(mn)):
            if not mn[i] < mx[i]:
                raise BoundaryError("Min must be smaller than Max for domain boundaries")
            n = math.floor((mx[i] - mn[i]) / interaction_range)
            m = to_usize(float(n))
            n_voxels.append(max(m, 1))
            dx.append((mx[i] - mn[i]) / float(n))
        return CartesianCuboid(mn, mx, n_voxels, dx)

    @staticmethod
    def from_boundaries_and_n_voxels(mn, mx, n_voxels):
        """:226-256"""
        dx = [(mx[i] - mn[i]) / float(n_voxels[i]) for i in range(len(mn))]
        return CartesianCuboid(mn, mx, n_voxels, dx)

    def get_all_voxel_indices(self):
        """:260-270 itertools multi_cartesian_product: the LAST axis varies fastest"""
        out = [[]]
        for i in range(self.D):
            out = [p + [k] for p in out for k in range(self.n_voxels[i])]
        return [tuple(p) for p in out]

    def get_voxel_index_of_raw(self, pos):
        """:320-337 (setup): check_min_max(min, pos) fails when min >= pos, then trunc((pos - min) / dx)"""
        for i in range(self.D):
            if self.min[i] >= pos[i]:
                raise BoundaryError("Min must be smaller than Max for domain boundaries")
        out = []
        for i in range(self.D):
            n_vox = (pos[i] - self.min[i]) / self.dx[i]
            u = to_usize(n_vox)
            if u is None:
                raise BoundaryError("Cannot convert float to usize")
            out.append(u)
        return tuple(out)

    def get_index_of(self, pos):
        """CartesianSubDomain::get_index_of :469-490 (every step, through SortCells :361-372)"""
        out = []
        for i in range(self.D):
            n_vox = (pos[i] - self.min[i]) / self.dx[i]
            u = to_usize(n_vox)
            if u is None:
                raise BoundaryError("Cannot convert float to usize")
            out.append(u)
        return tuple(out)

    def get_neighbor_voxel_indices(self, voxel_index):
        """:641-662"""
        bounds = []
        for i in range(self.D):
            bounds.append((max(voxel_index[i] - 1, 0), min(voxel_index[i] + 2, self.n_voxels[i])))
        out = [[]]
        for i in range(self.D):
            out = [p + [k] for p in out for k in range(bounds[i][0], bounds[i][1])]
        return [tuple(p) for p in out if tuple(p) != tuple(voxel_index)]

    def create_subdomains(self, n_subdomains):
        """:500-583: contiguous chunks of the voxel enumeration"""
        indices = self.get_all_voxel_indices()
        res = get_decomp_res(len(indices), n_subdomains)
        if res is None:
            raise RuntimeError("DecomposeError: Could not find a suiting decomposition")
        n, _m, average_len = res
        switcher = n * average_len
        groups: dict[int, list] = {}
        order = []
        for i, idx in enumerate(indices):
            if i < switcher:
                key = i // average_len
            else:
                key = (i - switcher) // max(average_len - 1, 1) + n
            if key not in groups:
                groups[key] = []
                order.append(key)
            groups[key].append(idx)
        return [(k, groups[k]) for k in order]

    def apply_boundary(self, pos, vel):
        """CartesianSubDomain::apply_boundary :594-631 (domain_min / domain_max of the WHOLE domain)"""
        position, velocity = list(pos), list(vel)
        two = 1.0 + 1.0
        for i in range(self.D):
            if position[i] < self.min[i]:
                position[i] = two * self.min[i] - position[i]
                velocity[i] = abs(velocity[i])
            if position[i] > self.max[i]:
                position[i] = two * self.max[i] - position[i]
                velocity[i] = -abs(velocity[i])
        for p, dmin, dmax in zip(position, self.min, self.max):
            if p < dmin or p > dmax:
                raise BoundaryError(f"Particle is out of domain at position {pos}")
        return position, velocity


# --------------------------------------------------------------------------------------------------------
# chili backend (cellular_raza-core/src/backend/chili)
# --------------------------------------------------------------------------------------------------------
@dataclass
class Cell:
    key: int          # index of the agent in the caller's input (the tests' join key)
    mech: object
    inter: object


@dataclass
class Voxel:
    plain_index: int
    neighbors: list   # ascending plain indices (BTreeSet)
    cells: list = field(default_factory=list)  # [(Cell, AuxStorageMechanics)]

    def calculate_force_between_cells_external(self, ext_pos, ext_vel, ext_inf):
        """update_mechanics.rs:65-108"""
        one_half = 1.0 / (1.0 + 1.0)
        force = None
        for cell, aux in self.cells:
            f1, f2 = cell.inter.calculate_force_between(list(cell.mech.pos), cell.mech.velocity(), ext_pos, ext_vel, ext_inf)
            aux.add_force(xa(f1, one_half))
            if force is not None:
                force = xapy(f2, one_half, force)
            else:
                force = xa(f2, one_half)
        return force


class SubDomainBox:
    def __init__(self, index):
        self.index = index
        self.voxels: dict[int, Voxel] = {}   # iterated in ascending plain index (BTreeMap)
        self.inbox_pos: list = []
        self.inbox_force: list = []
        self.inbox_cell: list = []


class ChiliTranscript:
    """One simulation of aspects [Mechanics, Interaction] as run_simulation! generates it
    (cellular_raza-core-proc-macro/src/run_sim.rs:240-328, 380-403), ``determinism: true`` (:141)."""

    def __init__(self, domain: CartesianCuboid, n_threads: int, dt: float, solver_order: int = 2):
        self.domain, self.dt, self.order = domain, dt, solver_order
        self.D = domain.D
        # construct_simulation_runner (datastructures.rs:51-217) after Domain::decompose
        # (cellular_raza-concepts-derive/src/domain.rs:292-420)
        self.subdomains = domain.create_subdomains(n_threads)
        self.voxel_index_to_plain_index = {}
        for _, voxels in self.subdomains:
            for v in voxels:
                self.voxel_index_to_plain_index[v] = len(self.voxel_index_to_plain_index)
        self.plain_index_to_subdomain = {}
        self.boxes: list[SubDomainBox] = []
        for si, (_, voxels) in enumerate(self.subdomains):
            box = SubDomainBox(si)
            for v in voxels:
                p = self.voxel_index_to_plain_index[v]
                self.plain_index_to_subdomain[p] = si
                nb = sorted(self.voxel_index_to_plain_index[w] for w in domain.get_neighbor_voxel_indices(v))
                box.voxels[p] = Voxel(p, nb)
            self.boxes.append(box)
        self.iteration = 0
        self.last_migrations: list = []   # (key, old plain, new plain)

    # -- insert_initial_cells (datastructures.rs:278-299) via decompose: input order inside a subdomain ----
    def set_cells(self, cells: Sequence[Cell]):
        per_sub: dict[int, list] = {}
        for c in cells:
            vi = self.domain.get_voxel_index_of_raw(c.mech.pos)
            if vi not in self.voxel_index_to_plain_index:
                raise IndexError_("could not find voxel index")
            per_sub.setdefault(self.plain_index_to_subdomain[self.voxel_index_to_plain_index[vi]], []).append(c)
        for box in self.boxes:
            for c in per_sub.get(box.index, []):
                vi = self.domain.get_index_of(c.mech.pos)
                p = self.voxel_index_to_plain_index[vi]
                box.voxels[p].cells.append((c, AuxStorageMechanics(self.D, self.order)))

    # -- one iteration of the generated main loop ----------------------------------------------------------
    def step(self, noise_row: Optional[Callable[[int], Sequence[float]]] = None):
        self.force_steps()
        self.update_steps(noise_row)

    def force_steps(self):
        """step_1 .. step_3 of the loop body (run_sim.rs:240-275): current_force of every cell afterwards is what
        local_mechanics_update integrates"""
        one_half = 1.0 / (1.0 + 1.0)
        # step 1: update_mechanics_interaction_step_1 (update_mechanics.rs:153-238)
        for box in self.boxes:
            for voxel_index in sorted(box.voxels.keys()):
                vox = box.voxels[voxel_index]
                n_cells = len(vox.cells)
                for n in range(n_cells):
                    c1, aux1 = vox.cells[n]
                    p1 = list(c1.mech.pos)
                    v1 = c1.mech.velocity()
                    i1 = c1.inter.get_interaction_information()
                    for m in range(n + 1, n_cells):
                        c2, aux2 = vox.cells[m]
                        p2 = list(c2.mech.pos)
                        v2 = c2.mech.velocity()
                        i2 = c2.inter.get_interaction_information()
                        force1, force2 = c1.inter.calculate_force_between(p1, v1, p2, v2, i2)
                        aux1.add_force(xa(force1, one_half))
                        aux2.add_force(xa(force2, one_half))
                        force2, force1 = c2.inter.calculate_force_between(p2, v2, p1, v1, i1)
                        aux1.add_force(xa(force1, one_half))
                        aux2.add_force(xa(force2, one_half))
                    force = None
                    for neighbor_index in vox.neighbors:
                        nvox = box.voxels.get(neighbor_index)
                        if nvox is not None:
                            f = nvox.calculate_force_between_cells_external(p1, v1, i1)
                            if f is not None:
                                force = xapy(f, 1.0, force) if force is not None else f
                        else:
                            dst = self.boxes[self.plain_index_to_subdomain[neighbor_index]]
                            dst.inbox_pos.append(dict(index_sender=voxel_index, index_receiver=neighbor_index,
                                                      pos=list(p1), vel=list(v1), info=i1, cell_index_in_vector=n))
                    # `force` (the return force of LOCAL neighbour voxels) goes out of scope here: it is never
                    # added to the cell (update_mechanics.rs:206-234)
                    del force
        # sync; step 2: update_mechanics_interaction_step_2 (:285-340)
        for box in self.boxes:
            received = box.inbox_pos
            box.inbox_pos = []
            received.sort(key=lambda pi: pi["index_sender"])  # stable, determinism = true
            for pos_info in received:
                vox = box.voxels.get(pos_info["index_receiver"])
                if vox is None:
                    raise IndexError_("EngineError: Voxel of PosInformation can not be found in this thread")
                force = vox.calculate_force_between_cells_external(pos_info["pos"], pos_info["vel"], pos_info["info"])
                if force is not None:
                    dst = self.boxes[self.plain_index_to_subdomain[pos_info["index_sender"]]]
                    dst.inbox_force.append(dict(force=force, cell_index_in_vector=pos_info["cell_index_in_vector"],
                                                index_sender=pos_info["index_sender"]))
        # sync; step 3: update_mechanics_interaction_step_3 (:344-382)
        for box in self.boxes:
            received = box.inbox_force
            box.inbox_force = []
            received.sort(key=lambda fi: fi["index_sender"])
            for obt in received:
                vox = box.voxels[obt["index_sender"]]
                vox.cells[obt["cell_index_in_vector"]][1].add_force(obt["force"])

    def update_steps(self, noise_row: Optional[Callable[[int], Sequence[float]]] = None):
        """local cell funcs + step_4 + step_5 of the loop body (run_sim.rs:280-328)"""
        D = self.D
        # local cell funcs: run_local_cell_funcs(local_mechanics_update) (datastructures.rs:305-326)
        for box in self.boxes:
            for p in sorted(box.voxels.keys()):
                for cell, aux in box.voxels[p].cells:
                    z_row = noise_row(cell.key) if noise_row is not None else [0.0] * D
                    mechanics_update(self.order, cell, aux, self.dt, z_row)
        # step 4: apply_boundary (:387-406), sort_cells_in_voxels_step_1 (:418-477)
        self.last_migrations = []
        for box in self.boxes:
            for p in sorted(box.voxels.keys()):
                for cell, _ in box.voxels[p].cells:
                    pos, vel = self.domain.apply_boundary(cell.mech.pos, cell.mech.velocity())
                    cell.mech.pos = pos
                    cell.mech.set_velocity(vel)
        for box in self.boxes:
            find_new_home_cells = []
            for voxel_index in sorted(box.voxels.keys()):
                vox = box.voxels[voxel_index]
                new_voxel_cells, old_voxel_cells = [], []
                for c, a in vox.cells:
                    moves = False
                    try:
                        index = self.domain.get_index_of(c.mech.pos)
                        plain = self.voxel_index_to_plain_index.get(index)
                        if plain is not None:
                            moves = plain != voxel_index
                    except BoundaryError:
                        moves = False  # errors are collected and dropped (:430-455)
                    (new_voxel_cells if moves else old_voxel_cells).append((c, a))
                find_new_home_cells.extend((voxel_index, ca) for ca in new_voxel_cells)
                vox.cells = old_voxel_cells
            for voxel_index, (cell, aux) in find_new_home_cells:
                ind = self.domain.get_index_of(cell.mech.pos)
                cell_index = self.voxel_index_to_plain_index[ind]
                self.last_migrations.append((cell.key, voxel_index, cell_index))
                tgt = box.voxels.get(cell_index)
                if tgt is not None:
                    tgt.cells.append((cell, aux))
                else:
                    dst = self.boxes[self.plain_index_to_subdomain[cell_index]]
                    dst.inbox_cell.append((voxel_index, cell, aux))
        # sync; step 5: sort_cells_in_voxels_step_2 (:485-513)
        for box in self.boxes:
            received = box.inbox_cell
            box.inbox_cell = []
            received.sort(key=lambda sc: sc[0])
            for _, cell, aux in received:
                index = self.voxel_index_to_plain_index[self.domain.get_index_of(cell.mech.pos)]
                vox = box.voxels.get(index)
                if vox is None:
                    raise IndexError_("Cell was sent to subdomain which does not hold this index")
                vox.cells.append((cell, aux))
        self.iteration += 1

    # -- views -----------------------------------------------------------------------------------------------
    def cells_in_order(self):
        """(plain index, Cell, Aux) in ascending voxel, in-voxel order: the order save_cells walks
        (datastructures.rs:373-396)"""
        out = []
        for box in self.boxes:
            for p in sorted(box.voxels.keys()):
                for cell, aux in box.voxels[p].cells:
                    out.append((p, cell, aux))
        out.sort(key=lambda t: t[0])  # subdomains hold ascending, disjoint plain-index chunks: stable
        return out

    def forces(self):
        """current_force of every cell after steps 1-3 (AuxStorageMechanics::current_force as
        local_mechanics_update would take it), computed on a deep copy: {key: force}"""
        import copy

        twin = copy.deepcopy(self)
        twin.force_steps()
        return {cell.key: (list(aux.current_force), aux.abs_sum) for _, cell, aux in twin.cells_in_order()}
