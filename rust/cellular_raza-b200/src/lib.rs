//! B200 backend of cellular_raza: a sibling of `chili`
//! (cellular_raza-core/src/backend/mod.rs:41-56) whose per-step mechanics/interaction path runs as
//! hand-written sm_100a kernels behind the C ABI of `include/crb.h`.
//!
//! STATUS: this crate has NOT been compiled — the build image has no cargo/rustc.  It is written
//! against `include/crb.h` (ABI 7) and mirrors `cellular_raza_b200/backend.py`, which drives the same
//! symbols through ctypes and is what `tests/` exercise.  Expect small fixes at first compile.
//!
//! What it replaces in the reference:
//! * `construct_simulation_runner` (chili/datastructures.rs:51-217)  -> [`B200Runner::new`] + `set_agents`
//! * the generated main loop (cellular_raza-core-proc-macro/src/run_sim.rs:380-403)
//!   -> [`B200Runner::step`] between save events
//! * `save_cells` at `TimeEvent::PartialSave` (chili/datastructures.rs:373-396)
//!   -> [`B200Runner::snapshot_async`] / [`B200Runner::snapshot_wait`]
//! * `SimulationError` variants of this path (chili/errors.rs:73-121) <- `crb_status` codes
pub mod ffi;

use cellular_raza_building_blocks::{
    BoundLennardJones, Brownian2D, Brownian3D, CartesianCuboid, Langevin2D, Langevin3D, MiePotential, MorsePotential,
    NewtonDamped2D, NewtonDamped3D,
};
use cellular_raza_core::backend::chili::SimulationError;
use cellular_raza_concepts::{BoundaryError, CalcError, DivisionError, IndexError};
use std::ffi::CStr;

/// Mechanics building blocks of the closed device model set pack themselves into one parameter row
/// (layout: `crb_mechanics_kind` in include/crb.h).
pub trait DeviceMechanics<const D: usize> {
    const KIND: u32;
    fn pack(&self, row: &mut [f64]);
    fn position(&self) -> [f64; D];
    fn velocity(&self) -> [f64; D];
}

/// Interaction building blocks (layout: `crb_interaction_kind`).
pub trait DeviceInteraction {
    const KIND: u32;
    fn pack(&self, row: &mut [f64]);
    fn species(&self) -> u8 {
        0
    }
}

macro_rules! impl_newton {
    ($t:ty, $d:literal) => {
        impl DeviceMechanics<$d> for $t {
            // cell_building_blocks/mechanics.rs:158-173: dx = vel, dv = force/mass - damping*vel
            const KIND: u32 = ffi::CRB_MECH_NEWTON_DAMPED;
            fn pack(&self, r: &mut [f64]) {
                r[0] = self.damping_constant;
                r[1] = self.mass;
            }
            fn position(&self) -> [f64; $d] {
                self.pos.into()
            }
            fn velocity(&self) -> [f64; $d] {
                self.vel.into()
            }
        }
    };
}
impl_newton!(NewtonDamped2D, 2);
impl_newton!(NewtonDamped3D, 3);

macro_rules! impl_brownian {
    ($t:ty, $d:literal) => {
        impl DeviceMechanics<$d> for $t {
            // mechanics.rs:336-357: dx = (D/kT) force; random dpos = sqrt(2 D) wiener
            const KIND: u32 = ffi::CRB_MECH_BROWNIAN;
            fn pack(&self, r: &mut [f64]) {
                r[0] = self.diffusion_constant;
                r[1] = self.kb_temperature;
            }
            fn position(&self) -> [f64; $d] {
                self.pos.into()
            }
            fn velocity(&self) -> [f64; $d] {
                [0.0; $d]
            }
        }
    };
}
impl_brownian!(Brownian2D, 2);
impl_brownian!(Brownian3D, 3);

macro_rules! impl_langevin {
    ($t:ty, $d:literal) => {
        impl DeviceMechanics<$d> for $t {
            // mechanics.rs:434-463
            const KIND: u32 = ffi::CRB_MECH_LANGEVIN;
            fn pack(&self, r: &mut [f64]) {
                r[0] = self.mass;
                r[1] = self.damping;
                r[2] = self.kb_temperature;
            }
            fn position(&self) -> [f64; $d] {
                self.pos.into()
            }
            fn velocity(&self) -> [f64; $d] {
                self.vel.into()
            }
        }
    };
}
impl_langevin!(Langevin2D, 2);
impl_langevin!(Langevin3D, 3);

impl DeviceInteraction for MorsePotential {
    // cell_building_blocks/interaction.rs:151-182, 226-260
    const KIND: u32 = ffi::CRB_INT_MORSE;
    fn pack(&self, r: &mut [f64]) {
        r[0] = self.radius;
        r[1] = self.potential_stiffness;
        r[2] = self.cutoff;
        r[3] = self.strength;
    }
}
impl DeviceInteraction for BoundLennardJones {
    // interaction.rs:122-138
    const KIND: u32 = ffi::CRB_INT_BOUND_LJ;
    fn pack(&self, r: &mut [f64]) {
        r[0] = self.epsilon;
        r[1] = self.sigma;
        r[2] = self.bound;
        r[3] = self.cutoff;
    }
}
impl DeviceInteraction for MiePotential {
    // interaction.rs:345-389
    const KIND: u32 = ffi::CRB_INT_MIE;
    fn pack(&self, r: &mut [f64]) {
        r[0] = self.radius;
        r[1] = self.strength;
        r[2] = self.bound;
        r[3] = self.cutoff;
        r[4] = self.en;
        r[5] = self.em;
    }
}

/// The species potential of cellular_raza-examples/cell_sorting/src/main.rs:24-97 (an example type, not a
/// building block): implement [`DeviceInteraction`] for it with `KIND = ffi::CRB_INT_CELL_SORTING`,
/// row `[cell_radius, potential_strength, relative_interaction_range]` and `species()` = the enum tag.
pub struct CellSortingRow {
    pub cell_radius: f64,
    pub potential_strength: f64,
    pub relative_interaction_range: f64,
    pub species: u8,
}
impl DeviceInteraction for CellSortingRow {
    const KIND: u32 = ffi::CRB_INT_CELL_SORTING;
    fn pack(&self, r: &mut [f64]) {
        r[0] = self.cell_radius;
        r[1] = self.potential_strength;
        r[2] = self.relative_interaction_range;
    }
    fn species(&self) -> u8 {
        self.species
    }
}

/// An agent of the device model set: one mechanics and one interaction building block.
pub struct DeviceAgent<M, I> {
    pub mechanics: M,
    pub interaction: I,
}

/// Settings that do not live in `chili::Settings` but are kwargs of `run_simulation!`
/// (cellular_raza-core-proc-macro/src/run_sim.rs:131-159) or specific to this backend.
#[derive(Clone, Copy, Debug)]
pub struct B200Options {
    /// `mechanics_solver_order` kwarg: N of MechanicsAdamsBashforthSolver<N> (default 2)
    pub mechanics_solver_order: u32,
    /// `CRB_WEIGHT_CHILI` reproduces chili's pair weights for `settings.n_threads` subdomains
    pub pair_weight_mode: u32,
    pub noise_mode: u32,
    pub device: u32,
}
impl Default for B200Options {
    fn default() -> Self {
        Self { mechanics_solver_order: 2, pair_weight_mode: ffi::CRB_WEIGHT_CHILI, noise_mode: ffi::CRB_NOISE_PHILOX, device: 0 }
    }
}

/// Owner of one `crb_sim` handle (all device memory lives behind it).  Not `Sync`: one host thread per
/// handle, as include/crb.h requires.
pub struct B200Runner<const D: usize> {
    sim: *mut ffi::crb_sim,
    n_mech: usize,
    n_inter: usize,
}

impl<const D: usize> Drop for B200Runner<D> {
    // cara/mod.rs:16-22 calls drop_runner the same way
    fn drop(&mut self) {
        unsafe { ffi::crb_destroy(self.sim) }
    }
}

/// Host copy of one snapshot (device order of the iteration it was taken at).
#[derive(Default)]
pub struct Snapshot<const D: usize> {
    pub pos: Vec<[f64; D]>,
    pub vel: Vec<[f64; D]>,
    pub ids: Vec<u64>,
    pub voxel: Vec<u64>,
    /// AuxStorageMechanics ring buffers, oldest first (chili/aux_storage.rs:41-130)
    pub hist_dx: Vec<[[f64; D]; 2]>,
    pub hist_dv: Vec<[[f64; D]; 2]>,
    pub hist_len: Vec<u32>,
}

impl<const D: usize> Snapshot<D> {
    /// AuxStorage of agent `k`: the ring buffers of increments, oldest first (chili/aux_storage.rs:41-130)
    pub fn aux(&self, k: usize) -> B200AuxStorage<D> {
        let mut m = cellular_raza_core::backend::chili::AuxStorageMechanics::default();
        use cellular_raza_core::backend::chili::UpdateMechanics;
        for e in 0..self.hist_len[k] as usize {
            m.set_last_position(self.hist_dx[k][e].into());
            m.set_last_velocity(self.hist_dv[k][e].into());
        }
        B200AuxStorage { mechanics: m }
    }
}

impl<const D: usize> B200Runner<D> {
    fn error(&self, code: i32) -> SimulationError {
        let msg = unsafe { CStr::from_ptr(ffi::crb_last_error(self.sim)) }.to_string_lossy().into_owned();
        match code {
            ffi::CRB_ERR_CALC => SimulationError::CalcError(CalcError(msg)),
            ffi::CRB_ERR_BOUNDARY => SimulationError::BoundaryError(BoundaryError(msg)),
            ffi::CRB_ERR_DIVISION => SimulationError::DivisionError(DivisionError(msg)),
            _ => SimulationError::IndexError(IndexError(format!("crb status {code}: {msg}"))),
        }
    }
    fn check(&self, code: i32) -> Result<(), SimulationError> {
        if code == ffi::CRB_OK {
            Ok(())
        } else {
            Err(self.error(code))
        }
    }

    /// `domain` supplies min / max / dx / n_voxels / rng_seed exactly as chili would use them
    /// (cartesian_cuboid_n.rs:183-256); `n_threads` is chili's subdomain count whose pair weights are
    /// reproduced (it creates no threads here).
    pub fn new<M: DeviceMechanics<D>, I: DeviceInteraction>(
        domain: &CartesianCuboid<f64, D>,
        dt: f64,
        n_threads: usize,
        capacity: usize,
        opt: B200Options,
    ) -> Result<Self, SimulationError> {
        let mut cfg = ffi::crb_config { abi_version: ffi::CRB_ABI_VERSION, dim: D as u32, ..Default::default() };
        cfg.mechanics_kind = M::KIND;
        cfg.interaction_kind = I::KIND;
        cfg.solver_order = opt.mechanics_solver_order;
        cfg.pair_weight_mode = opt.pair_weight_mode;
        cfg.ref_n_subdomains = n_threads.max(1) as u32;
        cfg.noise_mode = if M::KIND == ffi::CRB_MECH_NEWTON_DAMPED { ffi::CRB_NOISE_NONE } else { opt.noise_mode };
        cfg.device = opt.device;
        cfg.world_size = 1;
        let (mn, mx, dx, nv) = (domain.get_min(), domain.get_max(), domain.get_dx(), domain.get_n_voxels());
        for d in 0..D {
            cfg.domain_min[d] = mn[d];
            cfg.domain_max[d] = mx[d];
            cfg.dx[d] = dx[d];
            cfg.n_voxels[d] = nv[d] as u64;
        }
        cfg.dt = dt;
        cfg.rng_seed = domain.rng_seed;
        cfg.capacity = capacity as u64;
        let mut sim = std::ptr::null_mut();
        let rc = unsafe { ffi::crb_create(&cfg, &mut sim) };
        if rc != ffi::CRB_OK {
            let msg = unsafe { CStr::from_ptr(ffi::crb_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(SimulationError::IndexError(IndexError(format!("crb_create: status {rc}: {msg}"))));
        }
        Ok(Self {
            sim,
            n_mech: unsafe { ffi::crb_n_mech_params(M::KIND) } as usize,
            n_inter: unsafe { ffi::crb_n_inter_params(I::KIND) } as usize,
        })
    }

    /// Upload; identifiers are `CellIdentifier::Initial(k)` in input order like chili's
    /// (datastructures.rs:287-295 restarts per subdomain; here one global counter).
    pub fn set_agents<M: DeviceMechanics<D>, I: DeviceInteraction>(
        &mut self,
        agents: &[DeviceAgent<M, I>],
    ) -> Result<(), SimulationError> {
        let n = agents.len();
        let (mut pos, mut vel) = (Vec::with_capacity(n * D), Vec::with_capacity(n * D));
        let mut mech = vec![0.0f64; n * self.n_mech];
        let mut inter = vec![0.0f64; n * self.n_inter];
        let mut species = Vec::with_capacity(n);
        for (k, a) in agents.iter().enumerate() {
            pos.extend_from_slice(&a.mechanics.position());
            vel.extend_from_slice(&a.mechanics.velocity());
            a.mechanics.pack(&mut mech[k * self.n_mech..(k + 1) * self.n_mech]);
            a.interaction.pack(&mut inter[k * self.n_inter..(k + 1) * self.n_inter]);
            species.push(a.interaction.species());
        }
        let rc = unsafe {
            ffi::crb_set_agents(self.sim, n as u64, pos.as_ptr(), vel.as_ptr(), mech.as_ptr(), inter.as_ptr(),
                                species.as_ptr(), std::ptr::null())
        };
        self.check(rc)
    }

    /// Upload of user cells (packs through [`DeviceCell`]).
    pub fn set_cells<C: DeviceCell<D>>(&mut self, cells: &[C]) -> Result<(), SimulationError> {
        let n = cells.len();
        let (mut pos, mut vel) = (Vec::with_capacity(n * D), Vec::with_capacity(n * D));
        let mut mech = vec![0.0f64; n * self.n_mech];
        let mut inter = vec![0.0f64; n * self.n_inter];
        let mut species = Vec::with_capacity(n);
        for (k, c) in cells.iter().enumerate() {
            pos.extend_from_slice(&c.mechanics().position());
            vel.extend_from_slice(&c.mechanics().velocity());
            c.mechanics().pack(&mut mech[k * self.n_mech..(k + 1) * self.n_mech]);
            c.interaction().pack(&mut inter[k * self.n_inter..(k + 1) * self.n_inter]);
            species.push(c.interaction().species());
        }
        let rc = unsafe {
            ffi::crb_set_agents(self.sim, n as u64, pos.as_ptr(), vel.as_ptr(), mech.as_ptr(), inter.as_ptr(),
                                species.as_ptr(), std::ptr::null())
        };
        self.check(rc)
    }

    /// Replaces `n_steps` iterations of the generated loop body (run_sim.rs:380-403) for aspects
    /// [Mechanics, Interaction]; device-side CalcError / BoundaryError surface here.
    pub fn step(&mut self, n_steps: u64) -> Result<(), SimulationError> {
        let rc = unsafe { ffi::crb_step(self.sim, n_steps) };
        self.check(rc)
    }

    pub fn n_agents(&self) -> usize {
        unsafe { ffi::crb_n_agents(self.sim) as usize }
    }

    /// Starts a storage snapshot; `step` may be called while the device-to-host copy runs.
    pub fn snapshot_async(&mut self, snap: &mut Snapshot<D>) -> Result<(), SimulationError> {
        let n = self.n_agents();
        snap.pos.resize(n, [0.0; D]);
        snap.vel.resize(n, [0.0; D]);
        snap.ids.resize(n, 0);
        snap.voxel.resize(n, 0);
        snap.hist_dx.resize(n, [[0.0; D]; 2]);
        snap.hist_dv.resize(n, [[0.0; D]; 2]);
        snap.hist_len.resize(n, 0);
        let rc = unsafe {
            ffi::crb_snapshot_async(self.sim, n as u64, snap.pos.as_mut_ptr() as *mut f64, snap.vel.as_mut_ptr() as *mut f64,
                                    snap.ids.as_mut_ptr(), snap.voxel.as_mut_ptr(), snap.hist_dx.as_mut_ptr() as *mut f64,
                                    snap.hist_dv.as_mut_ptr() as *mut f64, snap.hist_len.as_mut_ptr())
        };
        self.check(rc)
    }

    /// Blocks until the snapshot started last is complete; `snap` must be the same buffers, untouched.
    pub fn snapshot_wait(&mut self) -> Result<usize, SimulationError> {
        let mut n = 0u64;
        let rc = unsafe { ffi::crb_snapshot_wait(self.sim, &mut n) };
        self.check(rc).map(|_| n as usize)
    }
}

/// A user agent of the device model set: what `#[derive(CellAgent)] struct Cell { #[Mechanics] mechanics,
/// #[Interaction] interaction }` is for chili (cellular_raza-examples/cell_sorting/src/main.rs:99-105).  The
/// backend needs to pack the agent into parameter rows and to write the stepped state back into a clone of it
/// for storage (`save_cells` serialises the user's own struct, chili/datastructures.rs:373-396).
pub trait DeviceCell<const D: usize>: Clone + serde::Serialize {
    type Mechanics: DeviceMechanics<D>;
    type Interaction: DeviceInteraction;
    fn mechanics(&self) -> &Self::Mechanics;
    fn interaction(&self) -> &Self::Interaction;
    /// `Position::set_pos` + `Velocity::set_velocity` of the concepts API
    fn set_state(&mut self, pos: &[f64; D], vel: &[f64; D]);
}

/// `_CrAuxStorage` of aspects `[Mechanics, Interaction]` as the proc-macro generates it
/// (cellular_raza-core-proc-macro/src/aux_storage.rs:807-823, 887-906): only Mechanics contributes a field.
#[derive(Clone, serde::Serialize, serde::Deserialize)]
pub struct B200AuxStorage<const D: usize> {
    pub mechanics: cellular_raza_core::backend::chili::AuxStorageMechanics<
        nalgebra::SVector<f64, D>,
        nalgebra::SVector<f64, D>,
        nalgebra::SVector<f64, D>,
        2,
    >,
}

/// The body of `run_main!` (cellular_raza-core-proc-macro/src/run_sim.rs:411-480) for this backend: storage
/// managers `<path>/cells` and `<path>/subdomains` (:432-444), the time stepper of `settings.time`
/// (cellular_raza-core/src/time.rs:173-266) driving batched `crb_step(n)` between `PartialSave` events, and
/// `StorageAccess { cells, subdomains }` as the result (chili/result.rs:5-10).
pub fn run_main<C, const D: usize>(
    domain: CartesianCuboid<f64, D>,
    agents: impl IntoIterator<Item = C>,
    settings: &cellular_raza_core::backend::chili::Settings<
        cellular_raza_core::time::FixedStepsize<f64>,
        false,
    >,
    opt: B200Options,
) -> Result<
    cellular_raza_core::backend::chili::StorageAccess<
        (cellular_raza_concepts::CellBox<C>, B200AuxStorage<D>),
        cellular_raza_building_blocks::CartesianSubDomain<f64, D>,
    >,
    SimulationError,
>
where
    C: DeviceCell<D> + for<'a> serde::Deserialize<'a>,
{
    use cellular_raza_concepts::{CellBox, CellIdentifier, Domain};
    use cellular_raza_core::storage::{StorageInterfaceStore, StorageManager};
    use cellular_raza_core::time::{TimeEvent, TimeStepper};

    let cells: Vec<C> = agents.into_iter().collect();
    let device: Vec<DeviceAgent<_, _>> = Vec::new();
    let _ = device;
    let n_threads: usize = settings.n_threads.into();
    let dt = settings.time.dt;
    let mut runner = B200Runner::<D>::new::<C::Mechanics, C::Interaction>(&domain, dt, n_threads, cells.len(), opt)?;
    runner.set_cells(&cells)?;

    // storage managers exactly as run_main builds them (run_sim.rs:432-444)
    let builder = settings.storage.clone().init();
    let mut storage_cells = StorageManager::<CellIdentifier, (CellBox<C>, B200AuxStorage<D>)>::open_or_create(
        builder.clone().suffix(builder.get_suffix().join("cells")), 0)?;
    let mut storage_subdomains = StorageManager::open_or_create(
        builder.clone().suffix(builder.get_suffix().join("subdomains")), 0)?;
    let subdomains = domain.create_subdomains(settings.n_threads)?;

    let mut time_stepper = settings.time.clone();
    let mut snap = Snapshot::<D>::default();
    let mut steps_since_save = 0u64;
    // save_cells + save_subdomains of one PartialSave event (chili/datastructures.rs:347-396)
    macro_rules! store_snapshot {
        ($iteration:expr) => {{
            let iteration: u64 = $iteration;
            runner.snapshot_async(&mut snap)?;
            runner.snapshot_wait()?;
            let batch = (0..snap.ids.len()).map(|k| {
                let id = CellIdentifier::Initial(snap.ids[k] as usize);
                let mut c = cells[snap.ids[k] as usize].clone();
                c.set_state(&snap.pos[k], &snap.vel[k]);
                (id, (CellBox::new(id, None, c), snap.aux(k)))
            });
            storage_cells.store_batch_elements(iteration, batch.collect::<Vec<_>>().iter().map(|(i, e)| (i, e)))?;
            for (index, subdomain, _) in subdomains.iter() {
                storage_subdomains.store_single_element(iteration, index, subdomain)?;
            }
        }};
    }
    // run_sim.rs:375-378: the stepper is at t0 when the loop starts, so the initial state is always stored
    if let Some(first) = time_stepper.save_initial() {
        store_snapshot!(first.iteration as u64);
    }
    while let Some(next) = time_stepper.advance()? {
        steps_since_save += 1;
        if let Some(TimeEvent::PartialSave) = next.event {
            runner.step(steps_since_save)?; // replaces that many iterations of run_sim.rs:380-403
            steps_since_save = 0;
            store_snapshot!(next.iteration as u64);
        }
    }
    if steps_since_save > 0 {
        runner.step(steps_since_save)?;
    }
    Ok(cellular_raza_core::backend::chili::StorageAccess { cells: storage_cells, subdomains: storage_subdomains })
}

/// `run_simulation!` of this backend: the same required keyword arguments as chili's macro
/// (cellular_raza-core-proc-macro/src/run_sim.rs:131-159, 594-607) and the same result type
/// `Result<StorageAccess<_, _>, SimulationError>` (chili/result.rs:5-10), so that a model switches backend by
/// changing the `use` line:
///
/// ```ignore
/// use cellular_raza_b200::run_simulation;           // was: cellular_raza::core::backend::chili::run_simulation
/// let storage = run_simulation!(
///     domain: domain,
///     agents: cells,
///     settings: settings,
///     aspects: [Mechanics, Interaction],
/// )?;
/// ```
///
/// Optional keyword arguments of the reference that this backend understands: `mechanics_solver_order: 0 | 1 | 2`
/// (:146), `determinism` (:141; the device path is always deterministic).  `parallelizer`, `zero_force_default`,
/// the `update_mechanics_interaction_step_*` plug-in points and `custom_update` belong to chili's generated loop and
/// are rejected at compile time.  `aspects` must be `[Mechanics, Interaction]`, `[Mechanics]` or those plus
/// `DomainForce` (anything else is outside the device model set).
#[macro_export]
macro_rules! run_simulation {
    (domain: $domain:expr, agents: $agents:expr, settings: $settings:expr, aspects: [$($asp:ident),+ $(,)?]
     $(, mechanics_solver_order: $order:expr)? $(, determinism: $det:expr)? $(,)?) => {{
        $( $crate::__check_aspect!($asp); )+
        #[allow(unused_mut)]
        let mut opt = $crate::B200Options::default();
        $( opt.mechanics_solver_order = $order; )?
        $( let _: bool = $det; )?
        $crate::run_main($domain, $agents, &$settings, opt)
    }};
}

#[doc(hidden)]
#[macro_export]
macro_rules! __check_aspect {
    (Mechanics) => {};
    (Interaction) => {};
    (DomainForce) => {};
    ($other:ident) => {
        compile_error!(concat!("aspect ", stringify!($other), " is outside the B200 backend's device model set"));
    };
}

/// The same loop as a plain function over already packed agents (no storage): `save` receives every snapshot.
/// `save` is called with every finished snapshot
/// (feed `StorageManager::store_batch_elements`, storage/concepts.rs:766-790); it runs while the device
/// already computes the next interval.
pub fn run_packed<M, I, const D: usize>(
    domain: &CartesianCuboid<f64, D>,
    agents: &[DeviceAgent<M, I>],
    dt: f64,
    n_threads: usize,
    save_iterations: &[u64],
    max_iterations: u64,
    opt: B200Options,
    mut save: impl FnMut(u64, &Snapshot<D>),
) -> Result<(), SimulationError>
where
    M: DeviceMechanics<D>,
    I: DeviceInteraction,
{
    let mut runner = B200Runner::<D>::new::<M, I>(domain, dt, n_threads, agents.len(), opt)?;
    runner.set_agents(agents)?;
    let mut snap = Snapshot {
        pos: vec![], vel: vec![], ids: vec![], voxel: vec![], hist_dx: vec![], hist_dv: vec![], hist_len: vec![],
    };
    let mut pending: Option<u64> = None;
    let mut it = 0u64;
    for &target in save_iterations.iter().filter(|&&s| s <= max_iterations) {
        if target > it {
            runner.step(target - it)?; // the copy of the previous snapshot overlaps this interval
            it = target;
        }
        if let Some(p) = pending.take() {
            runner.snapshot_wait()?;
            save(p, &snap);
        }
        runner.snapshot_async(&mut snap)?;
        pending = Some(it);
    }
    if it < max_iterations {
        runner.step(max_iterations - it)?;
    }
    if let Some(p) = pending.take() {
        runner.snapshot_wait()?;
        save(p, &snap);
    }
    Ok(())
}
