"""Drop-in replacements for the reference's ``network`` package (same import paths, class names,
constructor arguments, attribute names and state_dict layout); the eval-mode forward runs on the
native sm_100a kernels of ``libewvit.so``."""
