import sys, time, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import oraclepy as O, zpaq_b200 as Z
n=1044480; nblk=1029
pays=[O.gen_block(1,777+i,n) for i in range(nblk)]
ctx=Z.Context(0)
hdr=Z.model_from_method("5",pays[0])
t=time.time(); res=ctx.encode_blocks([hdr]*nblk,[b"\0"+p for p in pays]); tm=ctx.timing()
print("CHUNK",__import__('os').environ.get("ZPQ_CHUNK"),"ENC full", set(r[0] for r in res), tm, "MB/s %.1f"%(nblk*n/1e3/tm['code_ms']), flush=True)
