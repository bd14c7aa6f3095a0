/* gate_b200.h — C-ABI of the B200-native photon-in-voxel-volume dose engine.
 *
 * Plain C: pointers + sizes, no C++/torch/pybind types.  One engine handle is bound to one
 * CUDA device and is driven by one host thread.  All functions return 0 on success and a
 * non-zero GATE_B200_E* code on failure; gate_b200_last_error(h) gives the message.
 * The caller owns every host buffer it passes; the library owns device memory.
 *
 * The reference (OpenGATE/opengate @ 9b07951) has NO engine plug-in ABI: the
 * run phase is the pybind11 module `opengate_core` called from `opengate/engines.py`.
 * Each entry point below cites the reference interface whose job it takes over for the
 * photon-in-image-volume path; INTEGRATION.md shows the ctypes/pybind stub a maintainer
 * would add to `core/opengate_core`.
 *
 * Units: Geant4 internal (mm, MeV, ns); densities in g/cm3; cross-section tables in 1/mm.
 * Array layout: itk order, x fastest: flat = x + nx*(y + ny*z)
 * (reference core/opengate_core/opengate_lib/GateDoseActor.cpp:332-334).
 */
#ifndef GATE_B200_H
#define GATE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GATE_B200_ABI_VERSION 5

enum {
  GATE_B200_OK = 0,
  GATE_B200_EINVAL = 1,   /* bad argument / unsupported configuration */
  GATE_B200_ECUDA = 2,    /* CUDA runtime error (message has the cudaError string) */
  GATE_B200_ESTATE = 3,   /* call order violated (e.g. run before tables are set) */
  GATE_B200_ENOMEM = 4
};

typedef struct gate_b200_engine gate_b200_engine;

/* ---- physics tables -------------------------------------------------------------------
 * Replaces: Geant4 EM physics initialisation driven by opengate/engines.py:366-402
 * (PhysicsEngine.initialize_before/after_runmanager) and opengate/physics.py:300-337; the
 * table content follows the per-material export pattern of
 * GateMaterialMuHandler::SimulateMaterialTable (GateMaterialMuHandler.cpp:296-549).
 * Energy grid: n_energy nodes, log-uniform between e_min and e_max (inclusive).
 * Momentum-transfer grid: n_q nodes, log-uniform in q^2 (1/bohr^2) between q2_min, q2_max. */
typedef struct {
  int32_t n_materials;
  int32_t n_energy;
  double e_min_MeV, e_max_MeV;
  double e_cut_MeV;               /* photons below this are absorbed on the spot        */
  const float *density_g_cm3;     /* [n_materials]                                       */
  const float *mu_total_per_mm;   /* [n_materials][n_energy] total attenuation           */
  const float *cum_phot;          /* [n_materials][n_energy] P(phot)                     */
  const float *cum_compt;         /* [n_materials][n_energy] P(phot)+P(compt)            */
  const float *cum_rayl;          /* [n_materials][n_energy] ...+P(Rayl); rest = conv    */
  int32_t n_q;
  double q2_min, q2_max;
  const float *rayl_cum_ff2;      /* [n_materials][n_q] int_0^{q2} F_mol^2 d(q2)         */
  const float *compt_sf;          /* [n_materials][n_q] S_mol(q2)/S_mol(inf) in [0,1]    */
  int32_t pair_mode;              /* 0: whole E deposited locally (production-cut kerma
                                        configuration of the reference, apply_cuts=True),
                                     1: deposit E-2mc2 and emit two 511 keV photons      */
} gate_b200_tables;

/* ---- sources ----------------------------------------------------------------------------
 * Replaces: GateGenericSource::InitializePosition/Direction/Energy
 * (GateGenericSource.cpp:312-664), GateSPSEneDistribution.cpp:16-157,
 * GateSPSVoxelsPosDistribution.cpp:29-72, GateVSource.cpp:95-107 (attached-volume frame).
 * All frames are already composed to WORLD by the caller. */
enum { GATE_B200_POS_POINT = 0, GATE_B200_POS_BOX = 1, GATE_B200_POS_SPHERE = 2,
       GATE_B200_POS_DISC = 3, GATE_B200_POS_CYLINDER = 4, GATE_B200_POS_VOXELS = 5,
       GATE_B200_POS_PHSP = 6, GATE_B200_POS_SURFACE_SPHERE = 7 };
enum { GATE_B200_DIR_ISO = 0, GATE_B200_DIR_MOMENTUM = 1, GATE_B200_DIR_FOCUSED = 2,
       GATE_B200_DIR_PHSP = 3, GATE_B200_DIR_COS = 4, GATE_B200_DIR_BEAM2D = 5, GATE_B200_DIR_HISTOGRAM = 6 };
enum { GATE_B200_ENE_MONO = 0, GATE_B200_ENE_GAUSS = 1, GATE_B200_ENE_RANGE = 2,
       GATE_B200_ENE_SPECTRUM_DISCRETE = 3, GATE_B200_ENE_SPECTRUM_HISTOGRAM = 4,
       GATE_B200_ENE_SPECTRUM_HISTOGRAM_LINEAR = 5, GATE_B200_ENE_CDF = 6, GATE_B200_ENE_PHSP = 7 };

typedef struct {
  int32_t pos_type;
  double half_size[3];            /* box half sizes; cylinder: half_size[2] = dz         */
  double radius, sigma_x, sigma_y;
  double pos_rotation[9];         /* row-major, local -> world                            */
  double pos_translation[3];
  int32_t dir_type;
  double theta[2], phi[2];        /* iso: [min,max] in rad                                */
  double momentum[3];             /* world frame                                          */
  double focus_point[3];          /* world frame                                          */
  int32_t dir_relative;           /* 1: rotate sampled direction by dir_rotation          */
  double dir_rotation[9];
  int32_t ene_type;
  double mono_MeV, sigma_MeV, e_min_MeV, e_max_MeV;
  int32_t n_spectrum;             /* entries in spectrum_cdf                              */
  const double *spectrum_energy;  /* discrete: [n]; histogram: [n+1] edges; cdf: [n]      */
  const double *spectrum_cdf;     /* normalised cumulative weights, last == 1             */
  double weight;
  /* voxelised activity (pos_type == VOXELS): numpy (z,y,x) CDFs as built by
   * opengate/image.py:313-358 */
  int32_t vox_size[3];            /* nx, ny, nz                                           */
  double vox_spacing[3];
  double vox_origin[3];           /* centre of voxel (0,0,0) in the source frame          */
  const float *cdf_z;             /* [nz]                                                  */
  const float *cdf_y;             /* [nz][ny]                                              */
  const float *cdf_x;             /* [nz][ny][nx]                                          */
  /* phase-space replay (pos_type == PHSP, ene_type == PHSP, dir_type == PHSP or ISO).
   * Replaces: GatePhaseSpaceSource::GeneratePrimaries/GenerateOnePrimary
   * (GatePhaseSpaceSource.cpp:92-178) fed by PhaseSpaceSourceGenerator.generate
   * (sources/phspsources.py:119-258): the event with global id i replays entry
   * (phsp_entry_start + i) % phsp_n_entries (sequential reading, wrapping at the end of the
   * file).  Position: entry -> pos_rotation/pos_translation (identity for global_flag);
   * direction: entry, normalised, rotated by dir_rotation when dir_relative (dir_type ISO =
   * isotropic_direction); energy and weight (x `weight`) per entry.  Arrays are float32 as the
   * reference's batches (phspsources.py:195-210); phsp_weight may be NULL (= 1). */
  uint64_t phsp_n_entries;
  uint64_t phsp_entry_start;
  const float *phsp_pos_x, *phsp_pos_y, *phsp_pos_z;   /* [phsp_n_entries] mm            */
  const float *phsp_dir_x, *phsp_dir_y, *phsp_dir_z;   /* NULL with dir_type ISO         */
  const float *phsp_energy;                             /* MeV                            */
  const float *phsp_weight;
  /* dir_type BEAM2D (GateGenericSource.cpp:428-433 -> G4SPSAngDistribution beam2d): angular deviations
   * drawn from two Gaussians, rad.  dir_type COS (cosine law) uses theta / phi like ISO.  Both, and
   * pos_type SURFACE_SPHERE (points on the sphere of `radius`), are Geant4 G4SPS* algorithms that are
   * not in the reference tree: restated from the published sources. */
  double dir_sigma[2];
  /* dir_type HISTOGRAM (ABI 5; GateGenericSource.cpp:435-464 -> G4SPSAngDistribution "user", both angles): theta and
   * phi from user histograms -- bin i between edges[i] and edges[i+1] (rad) holds the probability cdf[i+1] - cdf[i]
   * (cdf[0] = 0, cdf[n] = 1: the weights are bin contents), the angle is uniform inside its bin (inverse of the
   * piecewise-linear cumulative, G4PhysicsFreeVector::GetEnergy); momentum = -(sin t cos p, sin t sin p, cos t). */
  int32_t n_dir_theta, n_dir_phi;                  /* bins */
  const double *dir_theta_edges, *dir_theta_cdf;   /* [n_dir_theta + 1] */
  const double *dir_phi_edges, *dir_phi_cdf;       /* [n_dir_phi + 1] */
  /* position.confine (ABI 5; GateGenericSource.cpp:376-383,223-227 -> G4SPSPosDistribution::ConfineSourceToVolume):
   * the position is drawn again until it lies inside the box below (a Box volume of the caller's geometry), at most
   * 100000 times as in Geant4 */
  int32_t confine;
  double confine_half_size[3];
  double world_to_confine[12];
} gate_b200_source;

/* ---- dose actor -------------------------------------------------------------------------
 * Replaces: GateDoseActor (GateDoseActor.cpp:144-243, 359-397) + AttachImageToVolume
 * (GateHelpersImage.txx:20-58). world_to_index maps a WORLD point p to the continuous
 * index c = A*p + b (row-major 3x4); the voxel is round-half-up(c) as itk's
 * TransformPhysicalPointToIndex does. */
enum { GATE_B200_HIT_RANDOM = 0, GATE_B200_HIT_PRE = 1, GATE_B200_HIT_POST = 2,
       GATE_B200_HIT_MIDDLE = 3 };
enum { GATE_B200_SCORE_EDEP = 1, GATE_B200_SCORE_EDEP_SQ = 2, GATE_B200_SCORE_DOSE = 4,
       GATE_B200_SCORE_DOSE_SQ = 8, GATE_B200_SCORE_COUNTS = 16 };
enum { GATE_B200_GRID_EDEP = 0, GATE_B200_GRID_EDEP_SQ = 1, GATE_B200_GRID_DOSE = 2,
       GATE_B200_GRID_DOSE_SQ = 3, GATE_B200_GRID_COUNTS = 4 };

typedef struct {
  int32_t size[3];
  double world_to_index[12];
  int32_t score_flags;            /* OR of GATE_B200_SCORE_*                              */
  int32_t hit_type;
} gate_b200_dose_desc;

/* ---- TLE dose actor ---------------------------------------------------------------------
 * Replaces: GateTLEDoseActor (GateTLEDoseActor.cpp:185-300): a DoseActor whose gammas score with
 * the track-length estimator.  Every step of a TLE gamma inside the attached image volume (one
 * step = the chord of one image voxel, as Geant4 steps through the nested parameterisation)
 * deposits  edep = w * 0.1 * E * (mu_en/rho)(material, E) * L[mm] * rho[g/cm3]  (:252-266) at a
 * point of the chord chosen by hit_type, dose = edep / rho; the secondaries of a TLE gamma do not
 * deposit (:60-74, :236-243).  TLE applies when E <= energy_threshold_MeV (tle_threshold_type
 * "energy", :224-226; +inf = type "None": always); above it the actor scores like a DoseActor.
 * A gamma at or below energy_min_MeV deposits w*E on the spot and stops (:268-271).
 * mu_en_over_rho is GateMaterialMuHandler::GetMuEnOverRho (GateMaterialMuHandler.cpp:75-79,
 * database "NIST" | "EPDL") evaluated by the caller on the energy nodes of gate_b200_tables.
 * The actor shares the id space and the grid accessors of the dose actors. */
typedef struct {
  double energy_min_MeV;
  double energy_threshold_MeV;
  const float *mu_en_over_rho;    /* [n_materials][n_energy] cm2/g                         */
} gate_b200_tle_desc;

/* ---- phase-space actor --------------------------------------------------------------------
 * Replaces: GatePhaseSpaceActor::SteppingAction (GatePhaseSpaceActor.cpp:119-201) for a Box
 * volume placed in the world, steps_to_store entering/exiting, and the digi attributes
 * KineticEnergy, PrePosition/Position, PreDirection/Direction, Weight, EventID,
 * GlobalTime/TimeFromBeginOfEvent (GateDigiAttributeList.cpp:42-279). */
enum { GATE_B200_PHSP_ENTERING = 1, GATE_B200_PHSP_EXITING = 2 };

typedef struct {
  double half_size[3];            /* of the Box the actor is attached to                  */
  double world_to_box[12];        /* row-major 3x4: p_box = A*p_world + b                 */
  int32_t steps_to_store;         /* OR of GATE_B200_PHSP_*                               */
  uint64_t capacity;              /* rows reserved on the device                          */
} gate_b200_phsp_desc;

typedef struct {
  double half_size[3];
  double world_to_box[12];        /* row-major 3x4: p_box = A*p_world + b                 */
  int32_t material;               /* index into the tables                                 */
} gate_b200_region_desc;

/* One stored row (structure of arrays on the device, gathered by gate_b200_get_phsp). */
typedef struct {
  float *kinetic_energy;          /* [n] MeV                                              */
  float *position;                /* [n][3] world, mm (PrePosition if entering)           */
  float *direction;               /* [n][3]                                               */
  float *weight;                  /* [n]                                                  */
  float *time_ns;                 /* [n] time from begin of event                         */
  int64_t *event_id;              /* [n]                                                  */
  int32_t *flags;                 /* [n] GATE_B200_PHSP_ENTERING / EXITING                */
} gate_b200_phsp_columns;

/* ---- run statistics -----------------------------------------------------------------------
 * Replaces: GateSimulationStatisticsActor counters (GateSimulationStatisticsActor.cpp:42-226)
 * plus the device counters the roofline figure is built from. */
typedef struct {
  uint64_t histories;             /* events                                               */
  uint64_t tracks;
  uint64_t steps;                 /* real interactions + boundary crossings              */
  uint64_t tentative_collisions;  /* Woodcock candidates = label gathers                  */
  uint64_t n_phot, n_compt, n_rayl, n_conv;
  uint64_t deposits;              /* deposit events before per-history coalescing        */
  uint64_t flushed_entries;       /* (voxel, history) pairs written to the grids         */
  uint64_t buffer_overflows;      /* per-history coalescing buffer evictions             */
  uint64_t escaped;               /* histories whose photon left the world               */
  uint64_t phsp_rows, phsp_dropped;
  uint64_t absorbed_events;       /* histories that never reached any phase-space plane  */
  double edep_total_MeV;          /* everything deposited (also outside scoring grids)   */
  double kernel_ms;               /* device time of the transport launches of this run   */
  uint64_t tle_steps;             /* voxel chords scored by track-length estimators      */
  uint64_t tle_entries;           /* TLE sums written to the grids (table entries/chords) */
  uint64_t tle_overflows;         /* per-history TLE table evictions (0 = squares exact) */
} gate_b200_stats;

/* ---- life cycle --------------------------------------------------------------------------- */
int gate_b200_abi_version(void);
/* Replaces: SimulationEngine.__init__/initialize_random_engine (engines.py:1414,1669-1698).
 * seed is the Philox key; history h uses counter (h, block) so results do not depend on the
 * number of GPUs or on scheduling. */
int gate_b200_create(int device_id, uint64_t seed, gate_b200_engine **out);
void gate_b200_destroy(gate_b200_engine *h);
/* Device memory of destroyed engines stays in the device's stream-ordered pool so that the next engine of
 * the process gets it without a trip to the driver (GATE_B200_POOL_RETAIN=0 disables this).  This call
 * hands the cached blocks back (cudaMemPoolTrimTo 0), e.g. before another library needs the memory. */
int gate_b200_release_cached_memory(int device_id);
const char *gate_b200_last_error(const gate_b200_engine *h);
/* Launch on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = own. */
int gate_b200_set_stream(gate_b200_engine *h, void *cuda_stream);

int gate_b200_set_tables(gate_b200_engine *h, const gate_b200_tables *t);
/* Replaces: VolumeEngine.Construct for the world box (engines.py:1166-1181). */
int gate_b200_set_world(gate_b200_engine *h, const double half_size[3], int32_t material);
/* Replaces: ImageVolume.construct + GateImageNestedParameterisation
 * (volumes.py:1049-1125, GateImageNestedParameterisation.cpp:19-75). labels: nx*ny*nz,
 * x fastest, label_bytes 1 or 2 (the reference uses unsigned short). world_to_volume maps a
 * world point to the volume frame whose origin is the volume centre. */
int gate_b200_set_image_volume(gate_b200_engine *h, const void *labels, int32_t label_bytes,
                               const int32_t size[3], const double spacing[3],
                               const double world_to_volume[12]);
/* Replaces: ImageVolume.create_label_image (geometry/volumes.py:1200-1243) + the above: the HU image goes
 * to the device and the label volume is made there, label = bin_labels[np.digitize(hu, bins)]
 * (bins sorted ascending, n_bins entries; bin_labels n_bins + 1 entries; integer-exact: the comparison
 * is done in double like numpy's).  labels_out (may be NULL) receives the label volume
 * (nx*ny*nz entries of label_bytes) for callers that want the host copy the reference keeps. */
enum { GATE_B200_HU_INT16 = 0, GATE_B200_HU_FLOAT32 = 1, GATE_B200_HU_FLOAT64 = 2 };
int gate_b200_set_image_volume_hu(gate_b200_engine *h, const void *hu, int32_t hu_type, const int32_t size[3],
                                  const double spacing[3], const double world_to_volume[12], int32_t n_bins,
                                  const double *bins, const uint16_t *bin_labels, int32_t label_bytes,
                                  void *labels_out);
int gate_b200_add_source(gate_b200_engine *h, const gate_b200_source *s, int32_t *source_id);
int gate_b200_add_dose_actor(gate_b200_engine *h, const gate_b200_dose_desc *d, int32_t *actor_id);
int gate_b200_add_tle_dose_actor(gate_b200_engine *h, const gate_b200_dose_desc *d,
                                 const gate_b200_tle_desc *t, int32_t *actor_id);
int gate_b200_add_phsp_actor(gate_b200_engine *h, const gate_b200_phsp_desc *d, int32_t *actor_id);

/* A homogeneous Box of its own material placed in the world (a couch, a detector slab, a collimator plate): what
 * the reference builds from a BoxVolume with `material` (geometry/volumes.py BoxVolume -> G4Box + G4LogicalVolume,
 * engines.py:924-933 for the actors of a volume).  Photons travelling in the world are attenuated and scatter in it
 * with its material's tables; its energy deposits count in the statistics only (dose actors are attached to the image
 * volume).  At most 4; the boxes must not overlap each other or the image volume (caller's responsibility).  ABI 5. */
int gate_b200_add_box_region(gate_b200_engine *h, const gate_b200_region_desc *d, int32_t *region_id);

/* Replaces: GateSourceManager::StartMasterThread run loop (GateSourceManager.cpp:198-244):
 * begin_run zeroes the accumulators (BeginOfRunActionMasterThread), run_source enqueues
 * n histories of one source whose global ids are first_history.. (EventID), sync waits. */
int gate_b200_begin_run(gate_b200_engine *h, int32_t run_id);
int gate_b200_run_source(gate_b200_engine *h, int32_t source_id, uint64_t first_history,
                         uint64_t n_histories);
int gate_b200_sync(gate_b200_engine *h);

/* Replaces: fetch_from_cpp_image (doseactors.py:255-270): copy one accumulator grid
 * (double, nx*ny*nz) to the host. */
int gate_b200_get_dose_grid(gate_b200_engine *h, int32_t actor_id, int32_t which, double *out);
/* The same with every value multiplied by `scale` on the device: the end-of-run unit conversion of
 * DoseActor.EndOfRunActionMasterThread (doseactors.py:685-696: dose /= Gy * voxel volume, squares by its square)
 * without a host pass over the image. */
int gate_b200_get_dose_grid_scaled(gate_b200_engine *h, int32_t actor_id, int32_t which, double scale, double *out);
/* The voxels [first, first + count) of one accumulator grid (x fastest), scaled: on a multi-GPU node every rank
 * holds the reduced grids after the all-reduce and copies a DIFFERENT slice into one host array shared by the
 * ranks, so that the device -> host read of the result uses the host links of all GPUs at once. */
int gate_b200_get_dose_grid_range(gate_b200_engine *h, int32_t actor_id, int32_t which, double scale,
                                  uint64_t first, uint64_t count, double *out);
/* Replaces: the uncertainty item of the reference's data containers (actors/dataitems.py:1956-1993 with
 * utility.py:551-560) evaluated on the device from the value and squared-value accumulators of one run:
 * out = sqrt(clip((S2/n - (S/n)^2) / (n - 1), 0)) / (S/n) where S != 0, else 0; which = GATE_B200_GRID_EDEP or
 * GATE_B200_GRID_DOSE (needs the matching squares); n = number of events (> 1). */
int gate_b200_get_uncertainty(gate_b200_engine *h, int32_t actor_id, int32_t which, uint64_t n_events, double *out);
/* Page-lock / unlock a caller-owned host buffer (cudaHostRegister): a get_* call into a registered buffer runs
 * at the full host-link rate, and the caller can prepare (allocate + register) its output arrays while a run is
 * in flight.  No reference counterpart (its images never leave host memory). */
int gate_b200_host_register(void *ptr, uint64_t bytes);
int gate_b200_host_unregister(void *ptr);
/* Device buffers that hold the accumulators of one dose actor, for the sum-reduce over GPUs done by
 * the host plumbing (NCCL; the reference merges parallel runs by summing images offline,
 * actors/dataitems.py:1238-1247).  On the device the scored ones of the sums edep, edep2, dose, dose2
 * of a voxel are the doubles of ONE record (stride 1, 2 or 4 doubles for 1, 2 or 3-4 scored sums, in
 * that order: a deposit touches one 32-byte DRAM sector): buffer 0 is that record array
 * (stride * nvox doubles), buffer 1 the counts image (nvox doubles) when it is scored.
 * get/set_dose_grid convert between the records and the dense x-fastest arrays of the ABI. */
int gate_b200_dose_buffers(gate_b200_engine *h, int32_t actor_id, int32_t max_buffers, void **ptrs,
                           uint64_t *n_elements, int32_t *n_buffers);
/* Host -> device overwrite of an accumulator (used after a reduce / for merging runs). */
int gate_b200_set_dose_grid(gate_b200_engine *h, int32_t actor_id, int32_t which, const double *in);
/* Replaces: GateDoseActor::ComputeMeanUncertainty (GateDoseActor.cpp:280-330). */
int gate_b200_mean_uncertainty(gate_b200_engine *h, int32_t actor_id, uint64_t n_events,
                               int32_t top_voxels, double threshold, double *out);
int gate_b200_phsp_count(gate_b200_engine *h, int32_t actor_id, uint64_t *n_rows);
int gate_b200_get_phsp(gate_b200_engine *h, int32_t actor_id, uint64_t max_rows,
                       const gate_b200_phsp_columns *out, uint64_t *n_rows);
int gate_b200_get_stats(gate_b200_engine *h, gate_b200_stats *out);
/* Name of the transport-kernel instantiation of the most recent launch of this engine, spelled as the CUDA
 * tools print it (e.g. "gb2::transport_kernel<unsigned char, 2>"; "" before the first launch): lets a caller tie
 * a measurement to the profiler capture of the SAME kernel variant (bench.py's roofline.traffic).  No reference
 * counterpart (the reference has no device code). */
const char *gate_b200_kernel_name(const gate_b200_engine *h);

/* One-call convenience used by benchmarks: upload nothing, just time `n` histories of an
 * already configured source (same as run_source + sync) and return device milliseconds. */
int gate_b200_timed_run(gate_b200_engine *h, int32_t source_id, uint64_t first_history,
                        uint64_t n_histories, double *device_ms);

#ifdef __cplusplus
}
#endif
#endif /* GATE_B200_H */
