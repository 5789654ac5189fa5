import sys, ctypes as C
sys.path.insert(0,'/root/repo/split-seq_demultiplexing_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, ssd_b200, fastq_util as fu
wl=fu.load_whitelist()
dm=ssd_b200.Demultiplexer(wl, errors=1, min_count=10, record_bytes_hint=160)
spec=ssd_b200.synth_spec(wl, 10_000_000)
d1,d2=ssd_b200.synth_device(dm, spec)
for i in range(3):
    dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
buf=(C.c_ulonglong*32)()
dm.lib.ssd_debug_scan_profile.argtypes=[C.c_void_p, C.c_int]
dm.lib.ssd_debug_scan_profile(buf, 1)
dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
dm.lib.ssd_debug_scan_profile(buf, 0)
names={0:'A.wait_full',1:'A.mask',2:'A.count+list',9:'A.wait_list_empty',5:'B.wait_list',6:'B.compute',3:'B.poll+sync',8:'B.poll_only',4:'B.commit',10:'A.popc+scan',11:'A.sync1',12:'A.bases+emit',13:'A.sync2',14:'B.header',15:'B.windows+umi'}
for m in (0,1):
    n=buf[m*16+7]
    print('R%d tiles=%d'%(m+1,n), {v: round(buf[m*16+k]/max(n,1)) for k,v in names.items()})
