import sys; sys.path.insert(0,'pynbody-extras_b200')
from pynbodyext.gravity import device as g
print({v: round(g.measure_fp32_peak(0,v),2) for v in range(4)})
