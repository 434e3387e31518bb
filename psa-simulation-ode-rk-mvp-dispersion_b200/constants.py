"""Physical constants (mirror of the reference's constants.py:2)."""
c = 299_792_458.0  # vacuum speed of light [m/s]
