"""Mirror of the reference's ``components`` package (same module, class and function names, same ``state_dict``
keys), dispatching to the sm_100a kernels.  ``garmentnets_b200.shims.install()`` aliases it as top-level
``components`` so the reference's ``networks/conv_implicit_wnf.py`` imports it unchanged."""
