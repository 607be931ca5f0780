// Command-line driver: same flags and console contract as the reference's remhos
// (remhos.cpp:216-340).  `-gpus N` (N > 1) plays the role of `mpirun -np N`: the
// driver starts one process per GPU (before anything touches CUDA), each of which finds its rank in the
// environment exactly as under `torchrun --no-python --nproc-per-node N remhos ...`.
#include "remhos_host.hpp"

#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <string>
#include <sys/wait.h>
#include <unistd.h>

static int run(int argc, char *argv[])
{
   double final_mass_u = 0.0;
   try
   {
      return remhos::remhos(argc, argv, final_mass_u);
   }
   catch (const std::exception &e)
   {
      // MFEM_VERIFY / MFEM_ABORT behaviour: message, then abnormal termination
      std::cerr << "\n\nVerification failed: " << e.what() << "\n ... in remhos_b200" << std::endl;
      return 134;   // the status a SIGABRT-terminated process reports
   }
}

int main(int argc, char *argv[])
{
   int gpus = 1;
   for (int i = 1; i + 1 < argc; i++)
   {
      if (!std::strcmp(argv[i], "-gpus") || !std::strcmp(argv[i], "--num-gpus")) { gpus = std::atoi(argv[i + 1]); }
   }
   if (gpus <= 1 || std::getenv("RANK")) { return run(argc, argv); }
   const std::string dir = "/dev/shm/rmh_rdzv_" + std::to_string((long)getpid());
   std::string rm = "rm -rf " + dir;
   if (std::system(rm.c_str()) != 0) { return 1; }
   setenv("RMH_RDZV_DIR", dir.c_str(), 1);
   setenv("WORLD_SIZE", std::to_string(gpus).c_str(), 1);
   for (int r = 0; r < gpus; r++)
   {
      const pid_t pid = fork();
      if (pid < 0) { std::cerr << "fork failed" << std::endl; return 1; }
      if (pid == 0)
      {
         setenv("RANK", std::to_string(r).c_str(), 1);
         setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
         _exit(run(argc, argv));
      }
   }
   int worst = 0;
   for (int r = 0; r < gpus; r++)
   {
      int st = 0;
      if (wait(&st) > 0)
      {
         const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
         if (rc > worst) { worst = rc; }
      }
   }
   if (std::system(rm.c_str()) != 0) { return worst ? worst : 1; }
   return worst;
}
