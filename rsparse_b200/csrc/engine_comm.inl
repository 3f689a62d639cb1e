// engine_comm.inl -- part of engine.cu (included there; not a standalone translation unit).
// ------------------------------------------------------------------------------------------------------
// 3. communicator (one process per GPU)
// ------------------------------------------------------------------------------------------------------
// NCCL is bound at run time, not link time: a host process may already carry a different libnccl.so.2 (PyTorch
// bundles its own), and single-GPU users need none at all.  dlopen() returns whichever copy is already loaded.
struct NcclApi {
  void* h = nullptr;
  decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&::ncclCommInitRank) CommInitRank = nullptr;
  decltype(&::ncclCommDestroy) CommDestroy = nullptr;
  decltype(&::ncclAllReduce) AllReduce = nullptr;
  decltype(&::ncclAllGather) AllGather = nullptr;
  decltype(&::ncclBroadcast) Broadcast = nullptr;
  decltype(&::ncclSend) Send = nullptr;
  decltype(&::ncclRecv) Recv = nullptr;
  decltype(&::ncclGroupStart) GroupStart = nullptr;
  decltype(&::ncclGroupEnd) GroupEnd = nullptr;
  decltype(&::ncclGetErrorString) GetErrorString = nullptr;
  int load() {
    if (h) return B200ALS_OK;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(B200ALS_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define B200ALS_SYM(name)                                                                     \
    name = reinterpret_cast<decltype(name)>(dlsym(h, "nccl" #name));                        \
    if (!name) return fail(B200ALS_ENCCL, "libnccl.so.2 lacks nccl" #name)
    B200ALS_SYM(GetUniqueId); B200ALS_SYM(CommInitRank); B200ALS_SYM(CommDestroy); B200ALS_SYM(AllReduce);
    B200ALS_SYM(AllGather); B200ALS_SYM(Broadcast); B200ALS_SYM(GroupStart); B200ALS_SYM(GroupEnd);
    B200ALS_SYM(GetErrorString); B200ALS_SYM(Send); B200ALS_SYM(Recv);
#undef B200ALS_SYM
    return B200ALS_OK;
  }
};
static NcclApi g_nccl;
struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
static Comm g_comm;

extern "C" int b200als_comm_unique_id(void* id_out) {
  static_assert(sizeof(ncclUniqueId) == B200ALS_UNIQUE_ID_BYTES, "ncclUniqueId size");
  if (!id_out) return fail(B200ALS_EINVAL, "null id");
  TRY(g_nccl.load());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return B200ALS_OK;
}
extern "C" int b200als_comm_init(const void* id, int rank, int world_size) {
  if (!id || world_size < 1 || rank < 0 || rank >= world_size) return fail(B200ALS_EINVAL, "bad communicator arguments");
  TRY(ctx().init());
  TRY(g_nccl.load());
  if (g_comm.comm) return fail(B200ALS_EINVAL, "communicator already initialised");
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  NC(g_nccl.CommInitRank(&g_comm.comm, world_size, uid, rank));
  g_comm.rank = rank;
  g_comm.world = world_size;
  return B200ALS_OK;
}
extern "C" int b200als_comm_destroy(void) {
  if (g_comm.comm) {
    g_nccl.CommDestroy(g_comm.comm);
    g_comm.comm = nullptr;
  }
  g_comm.rank = 0;
  g_comm.world = 1;
  return B200ALS_OK;
}
extern "C" int b200als_comm_info(int* rank, int* world_size) {
  if (rank) *rank = g_comm.rank;
  if (world_size) *world_size = g_comm.world;
  return B200ALS_OK;
}
