from oracle import splat as _os


def spherical_harmonics(degrees_to_use, viewdirs, coeffs, method="poly"):
    return _os.spherical_harmonics(degrees_to_use, viewdirs, coeffs)
