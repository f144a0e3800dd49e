// main.cpp -- SM_${NS}x${NT}: the reference's executable (src/main.cpp) on B200s.
//
// Same interactive protocol: ten parameters read from stdin by rank 0 with the prompts on stderr
// (ranks_x, ranks_t, m0, MD steps, trajectory length, beta, Ntherm, Nmeas, Nsteps, saveconf), the
// parameter banner and results on stdout, and the 2D_U1_{Nx}x{Nt}_m0{m0}_SimData.txt log.
// Instead of `mpirun -n N` the program forks ranks_x*ranks_t - 1 workers itself, one per GPU.
#include <chrono>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>

#include "b200_runtime.h"
#include "hmc.h"
#include "mpi_setup.h"

int main(int argc, char** argv) {
    (void)argc;
    (void)argv;
    int Ntherm = 0, Nmeas = 0, Nsteps = 0;
    double beta = 0, trajectory_length = 0, m0 = 0;
    int MD_steps = 0, saveconf = 0;

    CG::max_iter = 10000;
    CG::tol = 1e-10;

    std::cerr << "  -----------------------------" << std::endl;
    std::cerr << "|  Two-flavor Schwinger model   |" << std::endl;
    std::cerr << "| Hybrid Monte Carlo simulation |" << std::endl;
    std::cerr << "  -----------------------------" << std::endl;
    std::cerr << "Nx " << LV::Nx << " Nt " << LV::Nt << std::endl;
    std::cerr << "ranks_x: " << std::endl;
    std::cin >> mpi::ranks_x;
    std::cerr << "ranks_t: " << std::endl;
    std::cin >> mpi::ranks_t;
    std::cerr << "m0: " << std::endl;
    std::cin >> m0;
    std::cerr << "Molecular dynamics steps: " << std::endl;
    std::cin >> MD_steps;
    std::cerr << "Trajectory length: " << std::endl;
    std::cin >> trajectory_length;
    std::cerr << "beta: " << std::endl;
    std::cin >> beta;
    std::cerr << "Thermalization: " << std::endl;
    std::cin >> Ntherm;
    std::cerr << "Measurements: " << std::endl;
    std::cin >> Nmeas;
    std::cerr << "Step (sweeps between measurements): " << std::endl;
    std::cin >> Nsteps;
    std::cerr << "Save configurations yes/no (1 or 0): " << std::endl;
    std::cin >> saveconf;
    std::cerr << std::endl;
    if (!std::cin || mpi::ranks_x < 1 || mpi::ranks_t < 1) {
        std::cerr << "could not read the ten run parameters from stdin" << std::endl;
        return 1;
    }

    // one process per GPU; the workers inherit every parameter through fork (the reference
    // broadcasts them, main.cpp:60-69)
    b200::spawn_ranks(mpi::ranks_x * mpi::ranks_t);
    srand((mpi::rank + 1) * time(0));

    initializeMPI();            // tile widths, neighbour ranks, GPU context (+ NCCL)
    allocate_lattice_arrays();
    periodic_boundary();

    GaugeConf GConf = GaugeConf();

    std::string start_time_str;
    {
        const std::time_t now_c = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
        std::ostringstream tss;
        tss << std::put_time(std::localtime(&now_c), "%Y-%m-%d %H:%M:%S");
        start_time_str = tss.str();
    }
    const char* hostname = std::getenv("HOSTNAME");

    std::ostringstream NameData;
    NameData << "2D_U1_" << LV::Nx << "x" << LV::Nt << "_m0";
    {
        std::ostringstream m0_stream;
        m0_stream << std::setprecision(17) << m0;
        NameData << m0_stream.str();
    }
    NameData << "_SimData.txt";
    std::ofstream Datfile;
    if (mpi::rank == 0) {
        Datfile.open(NameData.str());
        Datfile << "#Date and time\n" << start_time_str << "\n";
        Datfile << "#Host\n" << (hostname ? hostname : "unknown") << "\n";
        Datfile << "#Nx      #Nt\n";
        Datfile << std::setw(10) << LV::Nx << std::setw(10) << LV::Nt << "\n";
        Datfile << "#ranks_x     #ranks_t     #ranks\n";
        Datfile << std::setw(15) << mpi::ranks_x << std::setw(15) << mpi::ranks_t << std::setw(15) << mpi::size << "\n";
        Datfile << "#beta                        #Ntherm     #Nmeas     #Nsteps\n";
        Datfile << std::setw(30) << std::setprecision(17) << beta << std::setw(11) << Ntherm << std::setw(11) << Nmeas
                << std::setw(11) << Nsteps << "\n";
        Datfile << "#trajectory_length     #MD_steps\n";
        Datfile << std::setw(30) << std::setprecision(17) << trajectory_length << std::setw(30) << MD_steps << "\n";
        Datfile << "#CG max iterations     #CG relative tolerance\n";
        Datfile << std::setw(30) << CG::max_iter << std::setw(30) << std::setprecision(17) << CG::tol << "\n";
        Datfile << "#m0\n";
        Datfile << std::setw(30) << std::setprecision(17) << m0 << "\n";
        Datfile.close();

        std::cout << "**********************************************************************" << std::endl;
        std::cout << "*                              PARAMETERS" << std::endl;
        std::cout << "* Nx = " << LV::Nx << ", Nt = " << LV::Nt << std::endl;
        std::cout << "* m0 = " << m0 << ", kappa = " << 1 / (2 * (m0 + 2)) << std::endl;
        std::cout << "* beta = " << beta << std::endl;
        std::cout << "* Thermalization confs = " << Ntherm << std::endl;
        std::cout << "* Measurement confs = " << Nmeas << std::endl;
        std::cout << "* Decorrelation steps (confs dropped between measurements) = " << Nsteps << std::endl;
        std::cout << "* Trajectory length = " << trajectory_length << ", Leapfrog steps = " << MD_steps
                  << ", Integration step = " << trajectory_length / MD_steps << std::endl;
        std::cout << "* CG max iterations = " << CG::max_iter << ", CG tolerance = " << CG::tol << std::endl;
        std::cout << "* Number of ranks on x = " << mpi::ranks_x << ", Number of ranks on t = " << mpi::ranks_t << std::endl;
        std::cout << "* Total number of MPI ranks = " << mpi::size << std::endl;
        std::cout << "* Each rank has " << mpi::maxSize << " lattice sites" << std::endl;
        std::cout << "* Host: " << (hostname ? hostname : "unknown") << std::endl;
        std::cout << "* Start time: " << start_time_str << std::endl;
        std::cout << "**********************************************************************" << std::endl;
    }

    HMC hmc = HMC(GConf, MD_steps, trajectory_length, Ntherm, Nmeas, Nsteps, beta, LV::Nx, LV::Nt, m0, saveconf);
    if (const char* start = std::getenv("SM_START_CONF")) {   // optional: resume from a .ctxt file
        GConf.readBinary(start);
        hmc = HMC(GConf, MD_steps, trajectory_length, Ntherm, Nmeas, Nsteps, beta, LV::Nx, LV::Nt, m0, saveconf);
        hmc.set_start_from_conf(true);
    }
    const double begin = b200::wtime();
    hmc.HMC_algorithm();
    const double end = b200::wtime();

    if (mpi::rank == 0) {
        const double elapsed_secs = end - begin;
        std::cout << "Average plaquette value / volume: Ep = " << hmc.getEp() << " dEp = " << hmc.getdEp() << std::endl;
        std::cout << "Average gauge action / volume: gS = " << hmc.getgS() << " dgS = " << hmc.getdgS() << std::endl;
        std::cout << "Acceptance rate: " << hmc.getacceptance_rate(Nmeas + Nsteps * Nmeas) << std::endl;
        std::cout << "Execution time = " << elapsed_secs << " s" << std::endl;
        std::cout << "-------------------------------" << std::endl;
        Datfile.open(NameData.str(), std::ios::app);
        Datfile << "#Ep                           #dEp\n";
        Datfile << std::setw(30) << std::setprecision(17) << hmc.getEp() << std::setw(30) << hmc.getdEp() << "\n";
        Datfile << "#gS                           #dgS\n";
        Datfile << std::setw(30) << std::setprecision(17) << hmc.getgS() << std::setw(30) << hmc.getdgS() << "\n";
        Datfile << "#Acceptance rate\n";
        Datfile << std::setw(30) << std::setprecision(17) << hmc.getacceptance_rate(Nmeas + Nsteps * (Nmeas - 1)) << "\n";
        Datfile << "#Execution time\n";
        Datfile << std::setw(30) << std::setprecision(17) << elapsed_secs;
        Datfile.close();

        // B200 work counters go to their own file so that _SimData.txt stays byte-compatible
        std::ostringstream jn;
        jn << "2D_U1_" << LV::Nx << "x" << LV::Nt << "_b200.json";
        std::ofstream js(jn.str());
        const double su = (double)hmc.dd_applications() * LV::Ntot;
        js << std::setprecision(12) << "{\"trajectories\": " << hmc.trajectories() << ", \"seconds\": " << elapsed_secs
           << ", \"traj_per_s\": " << hmc.trajectories() / elapsed_secs << ", \"device_seconds\": " << hmc.device_seconds()
           << ", \"dd_applications\": " << hmc.dd_applications() << ", \"dd_site_updates_per_s\": "
           << (hmc.device_seconds() > 0 ? su / hmc.device_seconds() : 0.0) << ", \"gpus\": " << mpi::size << "}\n";
    }

    free_lattice_arrays();
    b200::shutdown();
    return 0;
}
