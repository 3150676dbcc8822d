"""Per-layer conv_tc_kernel durations / TFLOP/s from an ncu launch list of `profile_image.py --net-only B` (GuidedResUnet)."""
import csv, sys
path, B = sys.argv[1], int(sys.argv[2])
lines=[l for l in open(path) if not l.startswith('==')]
rd=csv.DictReader(lines)
names=["c1.conv1","c1.conv2","pool1","c2.conv1","c2.conv2","pool2","c3.conv1","c3.conv2","pool3","c4.conv1","c4.conv2","pool4","c5.conv1","c5.conv2","upv6","sc6","c6.conv1","c6.conv2","upv7","sc7","c7.conv1","c7.conv2","upv8","sc8","c8.conv1","c8.conv2","upv9","sc9","c9.conv1","c9.conv2"]
FUSED = len(sys.argv) > 3 and sys.argv[3] == "fused"  # levels 8 and 9: transposed conv + 1x1 shortcut as one launch
def fl(mode,hw,cin,cout):
    if mode=='c': return 2*B*hw*hw*9*cin*cout
    if mode=='s2': return 2*B*(hw//2)**2*9*cin*cout
    if mode=='1': return 2*B*hw*hw*cin*cout
    if mode=='t': return 2*B*hw*hw*cin*4*cout
F=[fl('c',128,32,32)]*2+[fl('s2',128,32,64)]+[fl('c',64,64,64)]*2+[fl('s2',64,64,128)]+[fl('c',32,128,128)]*2+[fl('s2',32,128,256)]+[fl('c',16,256,256)]*2+[fl('s2',16,256,512)]+[fl('c',8,512,512)]*2
F+= [fl('t',8,512,256),fl('1',16,512,256)]+[fl('c',16,256,256)]*2+[fl('t',16,256,128),fl('1',32,256,128)]+[fl('c',32,128,128)]*2+[fl('t',32,128,64),fl('1',64,128,64)]+[fl('c',64,64,64)]*2+[fl('t',64,64,32),fl('1',128,64,32)]+[fl('c',128,32,32)]*2
if FUSED:
    for a, b in ((26, 27), (22, 23)):
        names[a:b + 1] = ["upsc" + names[a][-1]]
        F[a:b + 1] = [F[a] + F[b]]
tot=0; i=0; other=0
for r in rd:
    if r['Metric Name']!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',',''))/1e3
    if 'conv_tc' not in r['Kernel Name']:
        other+=v; continue
    tot+=v
    print(f"{names[i]:10s} {v:8.1f} us {F[i]/v/1e6:8.1f} TFLOP/s  time_share={0:4.1f}")
    i+=1
print(f"conv total {tot:.1f} us -> {sum(F)/tot/1e6:.1f} TFLOP/s; other kernels {other:.1f} us")
