import sys, ctypes as C
sys.path.insert(0,'/root/repo/split-seq_demultiplexing_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, ssd_b200, fastq_util as fu
wl=fu.load_whitelist()
dm=ssd_b200.Demultiplexer(wl, errors=1, min_count=10, record_bytes_hint=160)
spec=ssd_b200.synth_spec(wl, 10_000_000)
d1,d2=ssd_b200.synth_device(dm, spec)
for i in range(3):
    dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
buf=(C.c_ulonglong*16)()
dm.lib.ssd_debug_scan_profile.argtypes=[C.c_void_p, C.c_int]
dm.lib.ssd_debug_scan_profile(buf, 1)
dm.demux_device(d1.data_ptr(), d1.numel(), d2.data_ptr(), d2.numel())
dm.lib.ssd_debug_scan_profile(buf, 0)
names=['wait_full','mask','count+list','poll_prefix','records']
for m in (0,1):
    n=buf[m*8+7]
    print('R%d tiles=%d'%(m+1,n), {names[k]: round(buf[m*8+k]/max(n,1)) for k in range(5)}, 'cycles/tile total', round(sum(buf[m*8+k] for k in range(5))/max(n,1)))
