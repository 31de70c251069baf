"""stdin: output of ref_circle_test; prints how many of the 1000 probe points are classified
correctly (the figure tests/test_cpp_mirror.py::test_reference_circle_test_learns checks)."""
import re, sys
pts = re.findall(r"\(\((-?[\d.e+-]+),(-?[\d.e+-]+)\),(-?[\d.e+-]+)\)", sys.stdin.read())
print(len(pts), sum(((float(x) ** 2 + float(y) ** 2 <= 1.0) == (float(s) > 0.5)) for x, y, s in pts))
