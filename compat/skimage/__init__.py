"""stand-in for scikit-image: only io.imread (data/loveda.py:6)"""
