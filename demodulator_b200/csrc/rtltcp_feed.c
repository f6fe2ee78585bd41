/*
 * rtltcp_feed.c -- acquisition adapter: an rtl_tcp client that tunes the
 * dongle and hands the raw unsigned 8-bit IQ stream to the demodulator through a pipe.
 *
 *   rtltcp_feed [-a host] [-p port] [-f hz] [-s rate] [-g tenth_dB | -A] [-P ppm] [-T] [-c ctl] [-n bytes] \
 *       | demodulator -i - -o - -S48000 -l15000 ...
 *
 * Plays the part of the reference's rtltcp.py (/root/reference/rtltcp.py:71-90 ControlRtlTcp, :201-236
 * the command loop): connect to rtl_tcp, send the same opening commands (manual gain mode, AGC off,
 * bias tee off, sample rate), accept further "<command> <value>" pairs while the stream runs.  Where
 * rtltcp.py re-serves the samples to TCP clients (OutputServer, :106-198), this tool writes them to
 * stdout, because the consumer here is one process on the same host: the C host of the demodulator
 * treats a pipe as a live stream (one-chunk latency, slots cut short when the producer stalls).
 *
 * The rtl_tcp wire protocol (public, osmocom rtl-sdr): the server first sends 12 bytes -- "RTL0",
 * tuner type and gain count as big-endian u32 -- then raw u8 I,Q pairs; a command is 5 bytes, one
 * command byte and a big-endian u32 parameter.
 *
 * Further commands at run time: -c names a file or FIFO that is polled for lines "<command> <value>",
 * command = a number (0x01 ...) or one of the names below; a line "q" ends the feed.  No CUDA here; plain C99.
 */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <poll.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

static const struct {
    const char *name;
    unsigned char code;
} k_commands[] = {
    {"SET_FREQUENCY", 0x01}, {"SET_SAMPLE_RATE", 0x02}, {"SET_GAIN_MODE", 0x03}, {"SET_GAIN", 0x04},
    {"SET_FREQUENCY_CORRECTION", 0x05}, {"SET_IF_STAGE", 0x06}, {"SET_TEST_MODE", 0x07}, {"SET_AGC_MODE", 0x08},
    {"SET_DIRECT_SAMPLING", 0x09}, {"SET_OFFSET_TUNING", 0x0a}, {"SET_RTL_CRYSTAL", 0x0b}, {"SET_TUNER_CRYSTAL", 0x0c},
    {"SET_TUNER_GAIN_BY_INDEX", 0x0d}, {"SET_BIAS_TEE", 0x0e}, {"SET_TUNER_BANDWIDTH", 0x40},
    /* the extended set of the rtl_tcp forks, as the reference's helper lists it (rtltcp.py:32-63) */
    {"UDP_ESTABLISH", 0x41}, {"UDP_TERMINATE", 0x42}, {"SET_I2C_TUNER_REGISTER", 0x43}, {"SET_I2C_TUNER_OVERRIDE", 0x44},
    {"SET_TUNER_BW_IF_CENTER", 0x45}, {"SET_TUNER_IF_MODE", 0x46}, {"SET_SIDEBAND", 0x47}, {"REPORT_I2C_REGS", 0x48},
    {"GPIO_SET_OUTPUT_MODE", 0x49}, {"GPIO_SET_INPUT_MODE", 0x50}, {"GPIO_GET_IO_STATUS", 0x51}, {"GPIO_WRITE_PIN", 0x52},
    {"GPIO_READ_PIN", 0x53}, {"GPIO_GET_BYTE", 0x54}, {"IS_TUNER_PLL_LOCKED", 0x55}, {"SET_FREQ_HI32", 0x56},
};

static int send_all(int fd, const void *buf, size_t n) {
    const char *p = buf;
    while (n) {
        ssize_t w = send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += w;
        n -= (size_t) w;
    }
    return 0;
}

/* one command: a byte and a big-endian 32-bit parameter */
static int send_command(int sock, unsigned char cmd, uint32_t param) {
    unsigned char m[5] = {cmd, (unsigned char) (param >> 24), (unsigned char) (param >> 16), (unsigned char) (param >> 8),
                          (unsigned char) param};
    fprintf(stderr, "rtltcp_feed: command 0x%02x <- %u\n", cmd, param);
    return send_all(sock, m, sizeof m);
}

static int read_exact(int fd, void *buf, size_t n) {
    char *p = buf;
    while (n) {
        ssize_t r = read(fd, p, n);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return -1;
        p += r;
        n -= (size_t) r;
    }
    return 0;
}

static int write_all(int fd, const void *buf, size_t n) {
    const char *p = buf;
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += w;
        n -= (size_t) w;
    }
    return 0;
}

/* "<command> <value>": command by number or by name */
static int parse_line(const char *line, unsigned char *cmd, uint32_t *param) {
    char word[64];
    long long value;
    if (sscanf(line, "%63s %lli", word, &value) != 2) return -1;
    char *end;
    long code = strtol(word, &end, 0);
    if (*end == '\0' && code >= 0 && code <= 255) {
        *cmd = (unsigned char) code;
    } else {
        size_t i, n = sizeof k_commands / sizeof k_commands[0];
        for (i = 0; i < n; ++i)
            if (!strcasecmp(word, k_commands[i].name)) break;
        if (i == n) return -1;
        *cmd = k_commands[i].code;
    }
    *param = (uint32_t) value; /* negative values (ppm) travel as two's complement, like struct.pack('>Bi') */
    return 0;
}

int main(int argc, char **argv) {
    const char *host = "127.0.0.1", *port = "1234", *ctl_path = NULL;
    long long freq = -1, rate = 250000, gain = -1, ppm = 0, limit = -1;
    int agc = 0, bias_tee = 0, have_ppm = 0, opt;
    while ((opt = getopt(argc, argv, "a:p:f:s:g:AP:Tc:n:h")) != -1) {
        switch (opt) {
            case 'a': host = optarg; break;
            case 'p': port = optarg; break;
            case 'f': freq = (long long) strtod(optarg, NULL); break;
            case 's': rate = (long long) strtod(optarg, NULL); break;
            case 'g': gain = strtoll(optarg, NULL, 0); break;
            case 'A': agc = 1; break;
            case 'P': ppm = strtoll(optarg, NULL, 0); have_ppm = 1; break;
            case 'T': bias_tee = 1; break;
            case 'c': ctl_path = optarg; break;
            case 'n': limit = strtoll(optarg, NULL, 0); break;
            default:
                fprintf(stderr, "usage: rtltcp_feed [-a host] [-p port] [-f hz] [-s rate] [-g tenth_dB | -A] [-P ppm] [-T] "
                                "[-c control file or fifo] [-n bytes] > iq.u8\n");
                return opt == 'h' ? 0 : 2;
        }
    }

    struct addrinfo hints, *res = NULL;
    memset(&hints, 0, sizeof hints);
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    int gai = getaddrinfo(host, port, &hints, &res);
    if (gai) {
        fprintf(stderr, "rtltcp_feed: %s:%s: %s\n", host, port, gai_strerror(gai));
        return 1;
    }
    int sock = -1;
    for (struct addrinfo *a = res; a; a = a->ai_next) {
        sock = socket(a->ai_family, a->ai_socktype, a->ai_protocol);
        if (sock < 0) continue;
        if (!connect(sock, a->ai_addr, a->ai_addrlen)) break;
        close(sock);
        sock = -1;
    }
    freeaddrinfo(res);
    if (sock < 0) {
        perror("rtltcp_feed: connect");
        return 1;
    }
    {
        int sz = 8 << 20;
        setsockopt(sock, SOL_SOCKET, SO_RCVBUF, &sz, sizeof sz);
#ifdef F_SETPIPE_SZ
        fcntl(STDOUT_FILENO, F_SETPIPE_SZ, 1 << 20);
#endif
    }

    /* the dongle's greeting */
    unsigned char hdr[12];
    if (read_exact(sock, hdr, sizeof hdr) || memcmp(hdr, "RTL0", 4)) {
        fprintf(stderr, "rtltcp_feed: %s:%s did not greet like rtl_tcp\n", host, port);
        return 1;
    }
    fprintf(stderr, "rtltcp_feed: tuner type %u, %u gain steps\n",
            (unsigned) hdr[4] << 24 | (unsigned) hdr[5] << 16 | (unsigned) hdr[6] << 8 | hdr[7],
            (unsigned) hdr[8] << 24 | (unsigned) hdr[9] << 16 | (unsigned) hdr[10] << 8 | hdr[11]);

    /* opening commands: what the reference's helper sends on connect (manual gain mode unless AGC is asked
     * for, AGC mode, bias tee, sample rate), then what the command line asks for */
    int bad = 0;
    bad |= send_command(sock, 0x03, gain >= 0 || !agc ? 1u : 0u);
    bad |= send_command(sock, 0x08, agc ? 1u : 0u);
    bad |= send_command(sock, 0x0e, bias_tee ? 1u : 0u);
    bad |= send_command(sock, 0x02, (uint32_t) rate);
    if (freq >= 0) bad |= send_command(sock, 0x01, (uint32_t) freq);
    if (gain >= 0) bad |= send_command(sock, 0x04, (uint32_t) gain);
    if (have_ppm) bad |= send_command(sock, 0x05, (uint32_t) ppm);
    if (bad) {
        perror("rtltcp_feed: send");
        return 1;
    }

    int ctl = -1;
    if (ctl_path) {
        ctl = open(ctl_path, O_RDONLY | O_NONBLOCK);
        if (ctl < 0) perror("rtltcp_feed: control channel");
    }
    static unsigned char buf[1 << 20];
    char line[256];
    size_t line_len = 0;
    long long total = 0;
    int quit = 0;
    while (!quit) {
        struct pollfd pf[2] = {{sock, POLLIN, 0}, {ctl, POLLIN, 0}};
        if (poll(pf, ctl >= 0 ? 2 : 1, -1) < 0) {
            if (errno == EINTR) continue;
            break;
        }
        if (ctl >= 0 && (pf[1].revents & POLLIN)) {
            char ch;
            ssize_t r;
            while ((r = read(ctl, &ch, 1)) == 1) {
                if (ch == '\n' || line_len + 1 == sizeof line) {
                    line[line_len] = '\0';
                    unsigned char cmd;
                    uint32_t param;
                    if (line_len == 1 && (line[0] == 'q' || line[0] == 'Q')) quit = 1; /* as in the helper's loop (rtltcp.py:227) */
                    else if (line_len && !parse_line(line, &cmd, &param)) send_command(sock, cmd, param);
                    else if (line_len) fprintf(stderr, "rtltcp_feed: not a command: %s\n", line);
                    line_len = 0;
                } else {
                    line[line_len++] = ch;
                }
            }
        }
        if (ctl >= 0 && (pf[1].revents & (POLLHUP | POLLERR)) && !(pf[1].revents & POLLIN)) {
            /* the writer of a FIFO went away: reopen so that the next writer is heard */
            close(ctl);
            ctl = open(ctl_path, O_RDONLY | O_NONBLOCK);
        }
        if (pf[0].revents & (POLLIN | POLLHUP | POLLERR)) {
            size_t want = sizeof buf;
            if (limit >= 0 && (long long) want > limit - total) want = (size_t) (limit - total);
            if (!want) break;
            ssize_t r = read(sock, buf, want);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) break; /* the server closed the stream */
            if (write_all(STDOUT_FILENO, buf, (size_t) r)) break; /* the consumer went away */
            total += r;
        }
    }
    fprintf(stderr, "rtltcp_feed: %lld bytes of IQ forwarded\n", total);
    close(sock);
    return 0;
}
