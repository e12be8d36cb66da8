"""Test double for matplotlib (absent from this image): just enough of pyplot for the reference's post-processing scripts
(scripts/metrics.py, scripts/roc.py) to run UNCHANGED on our output files (tests/test_postprocessing.py). Nothing is drawn:
every plotted series is appended to the JSON-lines file named by $MPL_STUB_LOG so that the test can see what a figure would show."""
