"""The linear-Gaussian component of the reference's ``framework_extended`` on the same kernels."""
