// C entry points around the REFERENCE's own GateMuTable class and element database, for
// oracle/_ref/libref_mu.so.  The reference sources (GateMuTables.cpp, GateMuTables.h,
// GateMuDatabase.h) are compiled where they lie under /root/reference; this wrapper only loads the
// element tables the way GateMaterialMuHandler::InitElementTable does (GateMaterialMuHandler.cpp:265-295:
// raw triples per element) and exposes GateMuTable::GetMuOverRho / GetMuEnOverRho
// (GateMuTables.cpp:47-107).  Test infrastructure only.
#include <cmath>
#include <cstring>

#include "GateMuDatabase.h"
#include "GateMuTables.h"

extern "C" {

// number of nodes of element z in database db (0 = NIST, 1 = EPDL), -1 if out of range
int ref_mu_element_size(int db, int z) {
  const int n = db == 0 ? NIST_mu_muen_data_elementNumber : EPDL_mu_muen_data_elementNumber;
  const int *cnt = db == 0 ? NIST_mu_muen_data_energyNumber : EPDL_mu_muen_data_energyNumber;
  if (z < 0 || z > n) return -1;
  return cnt[z];
}

// raw nodes of element z: energy, mu/rho, mu_en/rho
int ref_mu_element_nodes(int db, int z, double *e, double *mu, double *mu_en) {
  const int size = ref_mu_element_size(db, z);
  if (size < 0) return -1;
  const int *cnt = db == 0 ? NIST_mu_muen_data_energyNumber : EPDL_mu_muen_data_energyNumber;
  const double *data = db == 0 ? NIST_mu_muen_data : EPDL_mu_muen_data;
  int index = 0;
  for (int i = 0; i < z; ++i) index += 3 * cnt[i];
  for (int j = 0; j < size; ++j) {
    e[j] = data[index]; mu[j] = data[index + 1]; mu_en[j] = data[index + 2];
    index += 3;
  }
  return size;
}

// Build a GateMuTable from n nodes given as natural logarithms (what ConstructMaterial stores) and
// evaluate the two look-ups at m energies.
void ref_mu_table_lookup(int n, const double *ln_e, const double *ln_mu, const double *ln_mu_en, double density,
                         int m, const double *energy, double *mu_over_rho, double *mu_en_over_rho,
                         double *mu_lin, double *mu_en_lin) {
  G4Material mat(density);
  G4MaterialCutsCouple couple(&mat);
  GateMuTable table(&couple, n);
  for (int i = 0; i < n; ++i) table.PutValue(i, ln_e[i], ln_mu[i], ln_mu_en[i]);
  for (int k = 0; k < m; ++k) {
    mu_over_rho[k] = table.GetMuOverRho(energy[k]);
    mu_en_over_rho[k] = table.GetMuEnOverRho(energy[k]);
    mu_lin[k] = table.GetMu(energy[k]);
    mu_en_lin[k] = table.GetMuEn(energy[k]);
  }
}

}  // extern "C"
