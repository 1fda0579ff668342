/* Headless stand-in for <ncurses.h>, used ONLY to compile the unmodified reference
 * translation unit (/root/reference/src/main.cpp) into oracle/_ref/ so that its
 * updatePhysics()/resolveCollisions() can be driven directly as the verbatim oracle.
 * TEST INFRASTRUCTURE - never part of the product path.
 *
 * Every entry point is a no-op; nothing here draws.  The list below is the set of
 * curses names the reference TU mentions.
 */
#ifndef TRACEIT_ORACLE_STUB_NCURSES_H
#define TRACEIT_ORACLE_STUB_NCURSES_H

typedef unsigned int chtype;
typedef unsigned long mmask_t;
typedef struct stub_window { int rows, cols; } WINDOW;
typedef struct { short id; int x, y, z; mmask_t bstate; } MEVENT;

static WINDOW stub_stdscr_storage = {50, 200};
static WINDOW* stdscr = &stub_stdscr_storage;
static int LINES = 50;
static int COLS = 200;

#ifndef OK
#define OK 0
#endif
#ifndef ERR
#define ERR (-1)
#endif
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define COLOR_BLACK 0
#define COLOR_RED 1
#define COLOR_GREEN 2
#define COLOR_YELLOW 3
#define COLOR_BLUE 4
#define COLOR_MAGENTA 5
#define COLOR_CYAN 6
#define COLOR_WHITE 7
#define COLOR_PAIR(n) ((chtype)(n) << 8)

#define ACS_VLINE ((chtype)'|')
#define ACS_HLINE ((chtype)'-')

#define KEY_DOWN 0402
#define KEY_UP 0403
#define KEY_LEFT 0404
#define KEY_RIGHT 0405
#define KEY_BACKSPACE 0407
#define KEY_ENTER 0527
#define KEY_MOUSE 0631

#define BUTTON4_PRESSED 0x00010000UL
#define BUTTON5_PRESSED 0x00200000UL
#define ALL_MOUSE_EVENTS 0x0fffffffUL
#define REPORT_MOUSE_POSITION 0x10000000UL

#define getmaxyx(win, y, x) do { (y) = (win) ? (win)->rows : 0; (x) = (win) ? (win)->cols : 0; } while (0)

static inline WINDOW* initscr(void) { return stdscr; }
static inline int endwin(void) { return OK; }
static inline int cbreak(void) { return OK; }
static inline int noecho(void) { return OK; }
static inline int nodelay(WINDOW*, bool) { return OK; }
static inline int keypad(WINDOW*, bool) { return OK; }
static inline int scrollok(WINDOW*, bool) { return OK; }
static inline int curs_set(int) { return OK; }
static inline bool has_colors(void) { return false; }
static inline int start_color(void) { return OK; }
static inline int init_pair(short, short, short) { return OK; }
static inline mmask_t mousemask(mmask_t m, mmask_t*) { return m; }
static inline int mouseinterval(int) { return 0; }
static inline int getmouse(MEVENT*) { return ERR; }

static inline WINDOW* newwin(int r, int c, int, int) { WINDOW* w = new WINDOW; w->rows = r; w->cols = c; return w; }
static inline int delwin(WINDOW* w) { if (w && w != stdscr) delete w; return OK; }
static inline int wbkgd(WINDOW*, chtype) { return OK; }
static inline int wattron(WINDOW*, int) { return OK; }
static inline int wattroff(WINDOW*, int) { return OK; }
static inline int box(WINDOW*, chtype, chtype) { return OK; }

static inline int getch(void) { return ERR; }
static inline int wgetch(WINDOW*) { return ERR; }
static inline int addch(chtype) { return OK; }
static inline int addstr(const char*) { return OK; }
static inline int mvwaddch(WINDOW*, int, int, chtype) { return OK; }
static inline int printw(const char*, ...) { return OK; }
static inline int mvprintw(int, int, const char*, ...) { return OK; }
static inline int mvwprintw(WINDOW*, int, int, const char*, ...) { return OK; }
static inline int move(int, int) { return OK; }
static inline int clear(void) { return OK; }
static inline int wclear(WINDOW*) { return OK; }
static inline int refresh(void) { return OK; }
static inline int wrefresh(WINDOW*) { return OK; }

#endif
