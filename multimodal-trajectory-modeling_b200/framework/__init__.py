"""Mirror of the reference's ``framework`` package for the EM / predict path."""
