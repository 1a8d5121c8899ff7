"""Dev check of the device transcript's `show`: decimal renderings of crafted coordinates, byte for byte."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproofspp_b200 as bp
from bulletproofspp_b200 import lib as L
ctx=bp.Context(0)
bad=0
vals=[1,5,10**9,10**9-1,2**32,2**64+1,10**18,2**200+12345,2**255+7,0,10**76,10**77+5]
for x in vals:
  for y in vals:
    t=C.c_void_p()
    ctx._ck(ctx.lib.bppp_dtr_create(ctx.h,1,20,1,C.byref(t)),"create")
    raw=L.int_to_le(x)+L.int_to_le(y)
    ctx._ck(ctx.lib.bppp_dtr_absorb(t,raw,1,1),"absorb")
    buf=C.create_string_buffer(4000); ln=C.c_size_t()
    ctx._ck(ctx.lib.bppp_dtr_export(t,0,buf,4000,C.byref(ln)),"export")
    dev=buf.raw[:ln.value]
    want=(str(x)+str(y)).encode()
    if dev!=want: bad+=1; print((x,y),"\n  dev ",dev,"\n  want",want)
    ctx.lib.bppp_dtr_destroy(t)
print("bad:",bad)
