/* transport_oracle.h — interface of the CPU oracle (test infrastructure only).
 * Takes the same descriptor structs as the product C-ABI (include/gate_b200.h) so tests
 * feed both sides identical inputs. */
#ifndef TRANSPORT_ORACLE_H
#define TRANSPORT_ORACLE_H
#include "../include/gate_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_engine oracle_engine;

typedef struct {
  float kinetic_energy;
  float position[3];
  float direction[3];
  float weight;
  float time_ns;
  int64_t event_id;
  int32_t flags;
} oracle_phsp_row;

#define ORACLE_TRACE_MAX 64
typedef struct {
  float E0, p0[3], d0[3];   /* primary */
  float E_end, edep;        /* photon energy when the history stopped, total deposited */
  int n_tentative, n_interactions, n_dep, n_iter;
  float dep_e[ORACLE_TRACE_MAX];
} oracle_trace;

oracle_engine *oracle_create(uint64_t seed);
void oracle_destroy(oracle_engine *g);
int oracle_set_tables(oracle_engine *g, const gate_b200_tables *t);
int oracle_set_world(oracle_engine *g, const double half[3], int material);
int oracle_set_image_volume(oracle_engine *g, const void *labels, int label_bytes, const int size[3],
                            const double spacing[3], const double w2v[12]);
int oracle_add_source(oracle_engine *g, const gate_b200_source *s);
int oracle_add_dose_actor(oracle_engine *g, const gate_b200_dose_desc *d);
int oracle_add_tle_dose_actor(oracle_engine *g, const gate_b200_dose_desc *d, const gate_b200_tle_desc *t);
int oracle_set_tle_fold(oracle_engine *g, int on);
int oracle_set_tle_rec_cap(oracle_engine *g, int cap);
int oracle_add_box_region(oracle_engine *g, const gate_b200_region_desc *d);
int oracle_add_phsp_actor(oracle_engine *g, const gate_b200_phsp_desc *d);
void oracle_begin_run(oracle_engine *g);
int oracle_run_source(oracle_engine *g, int source_id, uint64_t first, uint64_t n, int n_threads);
int oracle_trace_history(oracle_engine *g, int source_id, uint64_t hist, oracle_trace *out);
const double *oracle_dose_grid(oracle_engine *g, int actor, int which);
void oracle_get_stats(oracle_engine *g, gate_b200_stats *out);
uint64_t oracle_phsp_rows(oracle_engine *g, int actor, const oracle_phsp_row **rows);

float oracle_logf(float x);
float oracle_expf(float x);
void oracle_sincos2pi(float t, float *s, float *c);
void oracle_philox(uint64_t seed, uint64_t hist, uint32_t block, uint32_t out[4]);
void oracle_uniforms(uint64_t seed, uint64_t hist, int n, float *out);

#ifdef __cplusplus
}
#endif
#endif
