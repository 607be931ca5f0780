// Command-line driver: same flags and console contract as the reference's remhos
// (remhos.cpp:216-340).
#include "remhos_host.hpp"

#include <cstdlib>
#include <iomanip>
#include <iostream>

int main(int argc, char *argv[])
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
