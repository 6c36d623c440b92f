from oracle.thirdparty import magnitude_spectrum  # noqa: F401
