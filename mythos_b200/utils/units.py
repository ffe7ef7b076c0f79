"""oxDNA unit helpers (``mythos/utils/units.py:12-35``)."""


def get_kt(t_kelvin):
    """Temperature in Kelvin -> kT in simulation units."""
    return 0.1 * t_kelvin / 300.0


def get_kt_from_c(t_celsius):
    return get_kt(t_celsius + 273.15)


def from_kt(kt):
    return 300.0 * kt / 0.1
