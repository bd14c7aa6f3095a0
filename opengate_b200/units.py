"""Geant4 internal units as plain floats (mm, MeV, ns), same values as the reference's
hard-coded table (opengate/utility.py:103-...; Gy = 1e-12, g/cm3 = 6.2415e18)."""


class _Units(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


_e_SI = 1.602176634e-19
_joule = 1e-6 / _e_SI          # MeV
_s = 1e9                        # ns
_m = 1000.0                     # mm
_kg = _joule * _s * _s / (_m * _m)

g4_units = _Units(
    # length
    nm=1e-6, um=1e-3, mm=1.0, cm=10.0, m=1000.0, km=1e6,
    mm2=1.0, cm2=100.0, m2=1e6, mm3=1.0, cm3=1000.0, m3=1e9,
    # energy
    eV=1e-6, keV=1e-3, MeV=1.0, GeV=1e3, joule=_joule,
    # time
    ns=1.0, us=1e3, ms=1e6, s=1e9, second=1e9, min=60e9, hour=3600e9,
    # angle
    rad=1.0, mrad=1e-3, deg=0.017453292519943295, degree=0.017453292519943295,
    # mass / density
    kg=_kg, g=1e-3 * _kg, mg=1e-6 * _kg,
    g_cm3=1e-3 * _kg / 1000.0, mg_cm3=1e-6 * _kg / 1000.0, kg_m3=_kg / 1e9,
    # dose, activity
    Gy=_joule / _kg, gray=_joule / _kg,
    Bq=1.0 / _s, kBq=1e3 / _s, MBq=1e6 / _s, GBq=1e9 / _s,
    mole=1.0, g_mole=1e-3 * _kg,
)
# the reference also exposes the slash spellings used inside material database files
g4_units["g/cm3"] = g4_units["g_cm3"]
g4_units["mg/cm3"] = g4_units["mg_cm3"]
g4_units["kg/m3"] = g4_units["kg_m3"]
g4_units["g/mole"] = g4_units["g_mole"]
