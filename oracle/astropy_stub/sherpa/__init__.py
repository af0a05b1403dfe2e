"""Import-level stand-in for sherpa — test infrastructure only.  It lets the reference's
agnpy/fit/sherpa_wrapper.py be imported so that its module-level `_evaluate_sed_*_scenario` functions (the
parameter unpacking, log10 scaling, R_b and d_L derivation and component sum of the fit model) can be run
unmodified.  No fitting machinery."""
IS_AGNPY_B200_STUB = True
