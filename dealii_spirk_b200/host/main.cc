// Stand-alone driver with the reference's command line (main.cc:3608-3791): every argument is a JSON
// parameter file, run back to back, one statistics table at the end.  The dimension (the
// reference's compile-time IRK_DIMENSION, CMakeLists.txt:38-46) is the optional first argument
// "--dim=2|3" (default 3).  Single process / single GPU; multi-GPU runs are launched through
// `python -m dealii_spirk_b200.launch` (one process per GPU, NCCL id exchanged by torch.distributed).
#include <spirk_host.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

int main(int argc, char **argv)
{
  int dim = 3, first = 1;
  if (argc > 1 && std::strncmp(argv[1], "--dim=", 6) == 0)
    {
      dim   = std::atoi(argv[1] + 6);
      first = 2;
    }
  if (first >= argc)
    {
      std::fprintf(stderr, "usage: %s [--dim=2|3] input_0.json [input_1.json ...]\n", argv[0]);
      return 1;
    }
  std::cout << "Running in " << dim << "D on backend " << spirk_host_backend() << std::endl;
  for (int i = first; i < argc; ++i)
    {
      std::cout << std::endl << "Processing " << argv[i] << std::endl;
      spirk_run *run = nullptr;
      if (spirk_host_create(argv[i], 1, dim, 0, nullptr, 0, 1, 1, &run) || spirk_host_run(run))
        {
          std::cerr << std::endl
                    << "----------------------------------------------------" << std::endl
                    << "Exception on processing: " << std::endl
                    << spirk_host_last_error() << std::endl
                    << "Aborting!" << std::endl
                    << "----------------------------------------------------" << std::endl;
          spirk_host_destroy(run);
          return 1;
        }
      spirk_host_destroy(run);
    }
  return 0;
}
