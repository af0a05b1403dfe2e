"""CGS values of the astropy>=7 constants the reference uses (CODATA 2018 / IAU 2015);
agnpy/utils/conversion.py:1-9, agnpy/synchrotron/synchrotron.py:12.  `e`, `m_e`, `c` are pinned
bit-exactly by the reference's known answer nu_synch_peak(1 G, 100) = 27992489872.33304 Hz."""
c = 2.99792458e10
h = 6.62607015e-27
m_e = 9.1093837015e-28
m_p = 1.67262192369e-24
e = 4.803204712570263e-10
sigma_T = 6.6524587321e-25
G = 6.6743e-8
k_B = 1.380649e-16
sigma_sb = 5.670374419e-5   # astropy's (CODATA 2018) Stefan-Boltzmann constant in erg cm-2 s-1 K-4
M_sun = 1.988409870698051e33
mec2 = m_e * c**2
lambda_c_e = h / (m_e * c)
