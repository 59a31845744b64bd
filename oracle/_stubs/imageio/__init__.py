"""Import stub (test infrastructure): imageio is imported by src/visualization/plot_helper.py only."""
