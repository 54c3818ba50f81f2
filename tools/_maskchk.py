import torch, sys
sys.path.insert(0,'/root/repo')
from fcvsr_b200 import bands
from torchvision.transforms import Resize, functional as TF
import torchvision
print(torch.__version__, torchvision.__version__, torch.get_num_threads())
m1024 = bands._gaussian_masks_1024(4)
for (h,w) in ((36,40),(64,64),(180,320)):
    for nt in (1, 8, 16):
        torch.set_num_threads(nt)
        m = Resize([h,w], interpolation=TF.InterpolationMode.BICUBIC)(m1024)
        print(h,w,nt, float(m.double().sum()), float(m[1].abs().max()), float(m[3,h//2,w//2]))
