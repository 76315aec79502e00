V=dealii_spirk_b200/variants/libspirk_b200_minb1.so
B="python tools/bench_vmult.py --refine 6 --variants 0 --reps 20 --kernels apply cheb_step_own_dinv residual"
$B --tag base --nb 1
$B --tag grid148_pad --nb 1 --opt v3_grid=148 v3_smem_pad_kb=40
$B --tag grid148_nopad --nb 1 --opt v3_grid=148
$B --tag grid296 --nb 1 --opt v3_grid=296
$B --tag grid148_pad_npt2 --nb 1 --opt v3_grid=148 v3_smem_pad_kb=40 v3_npt=2
$B --tag grid148_pad_npt4 --nb 1 --opt v3_grid=148 v3_smem_pad_kb=40 v3_npt=4
$B --tag minb1_default --nb 1 --lib $V
$B --tag minb1_npt4 --nb 1 --lib $V --opt v3_npt=4
$B --tag minb1_npt2 --nb 1 --lib $V --opt v3_npt=2
$B --tag grid148_pad_nb2 --nb 2 --opt v3_grid=148 v3_smem_pad_kb=40
$B --tag minb1_nb2 --nb 2 --lib $V
$B --tag grid74_pad --nb 1 --opt v3_grid=74 v3_smem_pad_kb=40
$B --tag grid222 --nb 1 --opt v3_grid=222
