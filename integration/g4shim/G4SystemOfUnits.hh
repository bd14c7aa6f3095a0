/* TEST SHIM (see G4ShimTables.hh): the CLHEP units the exporter uses */
#pragma once
static const double mm = 1.0, cm = 10.0 * mm, cm3 = cm * cm * cm;
static const double MeV = 1.0;
static const double g = 6.241509074460763e+21;   /* CLHEP: gram in internal units (e+, MeV, mm, ns) */
