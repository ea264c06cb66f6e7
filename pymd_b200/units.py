"""Unit system of the reference (units.py:7-12): hbar = 1, energies in eV, time unit
0.658 fs, mass-scaled displacements; ``curcof`` converts mean(f.p) to nW."""
time = 0.658211814201041e-15
ohbar = 0.06466
hbar = 1.0
kb = 0.000086173423
length = ohbar
curcof = 243414.
