"""pymd_b200 -- B200-native ensemble implementation of pymd's generalized-Langevin
time-stepping path behind the reference's md / phbath / ebath class interfaces."""
__version__ = "0.1.0"
