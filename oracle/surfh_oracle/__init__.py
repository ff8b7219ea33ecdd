"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy fp64) of the reference's LMM
instrument operator, geometry and CG loop.  See model.py / instrument.py / thirdparty.py.
Never imported by the product package `surfh_b200`."""
